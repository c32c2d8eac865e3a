"""Host-side logic of the PauliObservable drop-in (no GPU): table build and cache format against the
reference's own tensors (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, HAM_CASES
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, synthetic
from anqs_quantum_chemistry_b200.pauli_observable import parse_of_qubit_operator_arrays


class _Op:
    def __init__(self, terms):
        self.terms = terms


@pytest.mark.parametrize('case', HAM_CASES)
def test_tables_match_reference(case, tmp_path):
    g = load_golden(case)
    n = int(g['qubit_num'])
    hs = HilbertSpace(qubit_num=n, device=torch.device('cpu'), parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(g['in_xy'], g['in_yz'], g['in_w'], n))
    assert ham.term_num == g['in_xy'].shape[0]
    assert ham.unq_xy_masks_num == g['unq_xy_masks'].shape[0]
    assert tuple(ham.unq_xy_masks.shape) == (ham.unq_xy_masks_num, 1)
    assert tuple(ham.rearranged_yz.shape) == (ham.term_num, 1)
    np.testing.assert_array_equal(ham.unq_xy_masks.numpy().reshape(-1), g['unq_xy_masks'])
    np.testing.assert_array_equal(ham.unq_xy_masks_inv.numpy(), g['unq_xy_masks_inv'])
    np.testing.assert_array_equal(ham.unq_xy_to_yz_num.numpy(), g['unq_xy_to_yz_num'])
    np.testing.assert_array_equal(ham.unq_xy_to_yz_start.numpy(), g['unq_xy_to_yz_start'])
    np.testing.assert_array_equal(ham.rearranged_yz.numpy().reshape(-1), g['rearranged_yz'])
    np.testing.assert_array_equal(ham.rearranged_weights.numpy(), g['rearranged_weights'])
    # the cache uses the reference's file names (PO:110-118) and is picked up by a second instance
    for name in ham.local_energy_structure_tensor_names:
        assert os.path.exists(os.path.join(str(tmp_path), f'{name}.npy'))
    ham2 = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(g['in_xy'], g['in_yz'], g['in_w'], n))
    np.testing.assert_array_equal(ham2.rearranged_weights.numpy(), g['rearranged_weights'])


@pytest.mark.parametrize('case', ['ham_n8_dense', 'ham_n64_sparse'])
def test_terms_dict_parse(case, tmp_path):
    g = load_golden(case)
    n = int(g['qubit_num'])
    terms = synthetic.pauli_arrays_to_terms(g['in_xy'].view(np.uint64), g['in_yz'].view(np.uint64), g['in_w'], n)
    w, xy, yz = parse_of_qubit_operator_arrays(_Op(terms), n)
    np.testing.assert_array_equal(xy, g['in_xy'])
    np.testing.assert_array_equal(yz, g['in_yz'])
    np.testing.assert_allclose(w, g['in_w'], atol=1e-15, rtol=0)
    hs = HilbertSpace(qubit_num=n, device=torch.device('cpu'), parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=_Op(terms))
    np.testing.assert_array_equal(ham.unq_xy_masks.numpy().reshape(-1), g['unq_xy_masks'])


def test_codec_matches_reference(tmp_path):
    g = load_golden('hilbert')
    hs20 = HilbertSpace(qubit_num=20, device=torch.device('cpu'), parent_dir=str(tmp_path), rng_seed=0)
    vec = hs20.base_idx2base_vec(torch.from_numpy(g['idx20']).view(-1, 1))
    np.testing.assert_array_equal(vec.numpy(), g['vec20'])
    np.testing.assert_array_equal(hs20.base_vec2base_idx(vec).numpy().reshape(-1), g['back20'])
    hs64 = HilbertSpace(qubit_num=64, device=torch.device('cpu'), parent_dir=str(tmp_path), rng_seed=0)
    vec = hs64.base_idx2base_vec(torch.from_numpy(g['a'][:32]).view(-1, 1))
    np.testing.assert_array_equal(vec.numpy(), g['vec64'])
    np.testing.assert_array_equal(hs64.base_vec2base_idx(vec).numpy().reshape(-1), g['back64'])
    # sort / unique are kernels (k2_sort.cu): no CPU path (their golden check is tests/test_gpu_sort.py)
    with pytest.raises(RuntimeError):
        hs64.sort_base_idx(torch.from_numpy(g['dup']).view(-1, 1))
    with pytest.raises(RuntimeError):
        hs64.compute_unique_indices(torch.from_numpy(g['dup']).view(-1, 1))


def test_compute_ops_refuse_cpu(tmp_path):
    hs = HilbertSpace(qubit_num=8, device=torch.device('cpu'), parent_dir=str(tmp_path), rng_seed=0)
    with pytest.raises(RuntimeError):
        hs.popcount(torch.zeros((4, 1), dtype=torch.int64))
    g = load_golden('ham_n8_dense')
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(g['in_xy'], g['in_yz'], g['in_w'], 8))
    with pytest.raises(RuntimeError):
        ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=torch.from_numpy(g['samples']).view(-1, 1),
                                           unq_batch_as_amps=torch.from_numpy(g['amps']), coupling_method='ham',
                                           alpha_num=2, beta_num=2)
