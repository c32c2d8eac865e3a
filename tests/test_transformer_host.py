"""CPU: the TransformerMADE mirror (construction order, parameter names, forward) against the logits produced by the
reference's own TransformerMADE module (tests/golden/tfm_*.npz), and the numpy masking oracle's basic invariants."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200.transformer_anqs import TransformerMADE
from oracle import anqs_numpy as onp

CASES = ['tfm_n12', 'tfm_n14', 'tfm_n20']


def build_net(g):
    torch.manual_seed(int(g['seed']))
    return TransformerMADE(dim=64, out_dim=4, depth=int(g['depth']), qubit_num=int(g['qubit_num']), head_num=int(g['head_num']),
                           dtype=torch.float64).eval()


@pytest.mark.parametrize('name', CASES)
def test_module_matches_reference(name):
    g = load_golden(name)
    net = build_net(g)
    assert [k for k, _ in net.named_parameters()] == list(g['param_names'])        # state_dicts interchange
    sums = np.array([[float(p.sum()), float((p * p).sum())] for p in net.parameters()])
    assert np.allclose(sums, g['param_checksums'], rtol=0, atol=1e-12)             # same seed -> same initial weights
    n = int(g['qubit_num'])
    x = g['samples'].view(np.uint64)
    bits = torch.from_numpy(((x[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int64))
    with torch.no_grad():
        assert np.abs(net(bits).numpy() - g['logits']).max() < 1e-12
        assert np.abs(net(bits[:, :n // 2]).numpy() - g['prefix_logits']).max() < 1e-12


@pytest.mark.parametrize('name', CASES)
def test_oracle_masking_invariants(name):
    g = load_golden(name)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    masks = onp.NumberSpinMasks(n, ne, qubit_per_qudit=1)
    assert masks.Q == n and masks.DM == 2
    x = g['samples'].view(np.uint64)
    lp = onp.transformer_log_psi_from_logits(g['logits'], x, masks)
    nphys = int(g['n_phys'])
    assert np.isfinite(lp[:nphys].real).all()
    ev = np.array([bin(int(v) & 0x5555555555555555).count('1') for v in x])
    od = np.array([bin(int(v) & 0xAAAAAAAAAAAAAAAA).count('1') for v in x])
    phys = (ev == ne // 2) & (od == ne // 2)
    assert np.array_equal(np.isneginf(lp.real), ~phys)
    # conditionals are normalised over the allowed outcomes at every level
    for t in (0, n // 2, n - 1):
        c = onp.transformer_cond_from_logits(g['logits'][:nphys], x[:nphys], t, masks)
        assert np.abs(np.exp(2 * c).sum(axis=1) - 1.0).max() < 1e-12
