"""GPU: the hand-written ordering kernels (k2_sort.cu) against torch.sort / torch.unique on the same inputs, bit for bit.
Reference semantics: HilbertSpace.sort_base_idx HS:239-261 (unsigned ascending, stable), compute_unique_indices HS:215-228
(signed ascending unique + inverse), the Gumbel sampler's sort(descending)[:k] ANQS:733."""
import numpy as np
import pytest
import torch

from anqs_quantum_chemistry_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0') if torch.cuda.is_available() else None


def _keys(n, bits, seed, dup=1):
    rng = np.random.default_rng(seed)
    hi = (1 << bits) - 1 if bits < 64 else (1 << 64) - 1
    k = rng.integers(0, hi, size=max(1, n // dup), dtype=np.uint64, endpoint=True)
    k = np.resize(k, n) if n else k[:0]
    rng.shuffle(k)
    return torch.from_numpy(k.view(np.int64)).to(DEV)


@pytest.mark.parametrize('n,bits,dup', [(0, 20, 1), (1, 20, 1), (31, 12, 1), (4097, 20, 3), (100003, 56, 1), (250000, 64, 2), (1 << 20, 36, 7)])
def test_sort_pairs_unsigned_stable(n, bits, dup):
    keys = _keys(n, bits, 1, dup)
    srt, perm = _lib.sort_pairs(keys, None, 0, bits)
    flipped = keys ^ (-(2 ** 63))  # unsigned order = signed order of key ^ 2^63
    _, ref_perm = torch.sort(flipped, stable=True)
    assert torch.equal(perm, ref_perm) and torch.equal(srt, keys[ref_perm])
    # explicit payloads travel with their keys
    vals = torch.arange(n, device=DEV, dtype=torch.int64) * 3 + 1
    srt2, v2 = _lib.sort_pairs(keys, vals, 0, bits)
    assert torch.equal(srt2, srt) and torch.equal(v2, vals[ref_perm])


@pytest.mark.parametrize('n,bits,dup', [(0, 14, 1), (5, 14, 1), (50000, 20, 4), (300000, 56, 3), (300000, 64, 3)])
def test_unique_matches_torch(n, bits, dup):
    keys = _keys(n, bits, 2, dup)
    unq, inv = _lib.unique_i64(keys, end_bit=bits)
    ref_u, ref_inv = torch.unique(keys, return_inverse=True)   # signed ascending
    assert torch.equal(unq, ref_u)
    if n:
        assert torch.equal(inv, ref_inv) and torch.equal(unq[inv], keys)


@pytest.mark.parametrize('n,k', [(1, 1), (100, 100), (5000, 17), (100000, 10000), (3000000, 10000), (3000000, 1), (200000, 150000)])
def test_topk_is_the_head_of_a_stable_descending_sort(n, k):
    g = torch.Generator(device='cpu').manual_seed(n + k)
    v = torch.randn(n, generator=g, dtype=torch.float64)
    v[torch.rand(n, generator=g) < 0.2] = float('-inf')          # masked children
    v[torch.rand(n, generator=g) < 0.1] = 0.25                     # ties (also across the cut)
    v = v.to(DEV)
    top_v, top_i = _lib.topk_f64(v, k)
    ref_v, ref_i = torch.sort(v, descending=True, stable=True)
    assert torch.equal(top_v, ref_v[:k]) and torch.equal(top_i, ref_i[:k])
    # unsorted variant: the same set, in position order
    set_v, set_i = _lib.topk_f64(v, k, sorted=False)
    assert torch.equal(set_i, torch.sort(ref_i[:k]).values) and torch.equal(set_v, v[set_i])


def test_descending_float_sort_with_negative_zero_and_infinities():
    v = torch.tensor([0.0, -0.0, 1.5, float('inf'), -2.0, float('-inf'), 1.5, -2.0, 1e-300, -1e-300], dtype=torch.float64, device=DEV)
    srt, perm = _lib.sort_pairs(v, None, 0, 64, key_kind=1, xor_mask=-1)
    assert srt.tolist()[:3] == [float('inf'), 1.5, 1.5] and perm.tolist()[1:3] == [2, 6] and srt.tolist()[-1] == float('-inf')
    assert (srt[:-1] >= srt[1:]).all()


def test_hilbert_space_methods_use_the_kernels(tmp_path):
    from anqs_quantum_chemistry_b200 import HilbertSpace
    hs = HilbertSpace(qubit_num=20, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    keys = _keys(30000, 20, 5, dup=2).view(-1, 1)
    unq, inv = hs.compute_unique_indices(keys)
    ref_u, ref_inv = torch.unique(keys[:, 0], return_inverse=True)
    assert torch.equal(unq[:, 0], ref_u) and torch.equal(inv, ref_inv)
    srt, perm = hs.sort_base_idx(keys)
    assert torch.equal(srt, keys[perm]) and bool((srt[1:, 0] >= srt[:-1, 0]).all())


def test_sort_and_unique_match_the_reference_golden(tmp_path):
    """HS:239-261 / HS:215-228 outputs of the unmodified reference (tests/golden/hilbert.npz), 64-qubit indices with bit 63 set."""
    from conftest import load_golden
    from anqs_quantum_chemistry_b200 import HilbertSpace
    g = load_golden('hilbert')
    hs64 = HilbertSpace(qubit_num=64, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    dup = torch.from_numpy(g['dup']).to(DEV).view(-1, 1)
    s, p = hs64.sort_base_idx(dup)
    np.testing.assert_array_equal(s.cpu().numpy().reshape(-1), g['sorted'])
    np.testing.assert_array_equal(p.cpu().numpy(), g['sort_perm'])
    if 'unq' in g:
        u, inv = hs64.compute_unique_indices(dup)
        np.testing.assert_array_equal(u.cpu().numpy().reshape(-1), g['unq'])
        np.testing.assert_array_equal(inv.cpu().numpy(), g['unq_inv'])


@pytest.mark.parametrize('name', ['ham_n8_dense', 'ham_n12_dense', 'ham_n20_dense', 'ham_n56_sparse', 'ham_n64_sparse'])
def test_gpu_table_builder_matches_reference_tables(name, tmp_path):
    """the six local-energy structure tensors built on the GPU (radix sort / unique / segment) equal
    the reference's (PO:131-211), from the arrays and from the `.terms` dictionary."""
    from conftest import load_golden
    from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator
    g = load_golden(name)
    n = int(g['qubit_num'])
    op = PauliArraysOperator(g['in_xy'], g['in_yz'], g['in_w'], n)

    class FromTerms:
        terms = op.terms
    for sub, operator in (('a', op), ('t', FromTerms())):
        hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=str(tmp_path / sub), rng_seed=0)
        ham = PauliObservable(hilbert_space=hs, of_qubit_operator=operator)
        for key in ham.local_energy_structure_tensor_names:
            ref = g[key].reshape(-1)
            got = getattr(ham, key).cpu().numpy().reshape(-1)
            assert got.shape == ref.shape, key
            if key == 'rearranged_weights':
                assert np.abs(got - ref).max() == 0.0
            else:
                assert np.array_equal(got, ref), key
