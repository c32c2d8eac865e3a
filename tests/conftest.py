import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'needs_reference: needs /root/reference (build container only); skipped elsewhere')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    have_ref = ref_shim.reference_available()
    for item in items:
        if 'needs_reference' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='reference tree not present'))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))


HAM_CASES = ['ham_n8_dense', 'ham_n12_dense', 'ham_n14_dense', 'ham_n20_dense', 'ham_n56_sparse', 'ham_n64_sparse']
HAM_CASES_WITH_LISTS = ['ham_n8_dense', 'ham_n12_dense', 'ham_n64_sparse']


def c5_full_inputs():
    """The headline Hamiltonian as oracle/make_golden.py:make_ham_c5 fed it to the reference: synthetic_hamiltonian(56 qubits, 8
    irreps, seed 0), term order shuffled by default_rng(7); checked against the checksums the golden keeps of the inputs."""
    from anqs_quantum_chemistry_b200 import synthetic
    g = load_golden('ham_c5_full')
    xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
    perm = np.random.default_rng(7).permutation(xy.shape[0])
    xy, yz, w = xy[perm], yz[perm], w[perm]
    cs = g['in_checksums']
    assert int(np.bitwise_xor.reduce(xy.view(np.int64))) == int(cs[0]) and int(np.bitwise_xor.reduce(yz.view(np.int64))) == int(cs[1])
    assert xy.shape[0] == int(cs[2]) and abs(w.sum().real - g['in_w_sum'][0]) < 1e-9 and abs((np.abs(w) ** 2).sum() - g['in_w_sum'][1]) < 1e-9
    return g, xy, yz, w
