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
