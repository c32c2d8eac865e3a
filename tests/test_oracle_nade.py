"""CPU: NADE mode.  The drop-in's parameter layout / initial weights and the numpy oracle against the golden vectors the
unmodified reference produced with de_mode='NADE' (tests/golden/nade_*.npz)."""
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         LogAbsPhaseANQS, ANQSConfig)
from oracle import anqs_numpy as onp

CASES = ['nade_n12', 'nade_n20', 'nade_n56']


def build(g, device='cpu'):
    n, ne, seed = int(g['qubit_num']), int(g['particle_num']), int(g['seed'])
    hs = HilbertSpace(qubit_num=n, device=device, parent_dir=tempfile.mkdtemp(prefix='anqs_nade_test_'), rng_seed=seed)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(seed)
    return LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='NADE'))


def numpy_weights(wf):
    def layers(nets):
        return ([[l.weight.detach().cpu().numpy() for l in m.layers] for m in nets],
                [[l.bias.detach().cpu().numpy() for l in m.layers] for m in nets])
    Wa, ba = layers(wf.log_abs_subnet)
    Wp, bp = layers(wf.phase_subnet)
    return Wa, ba, Wp, bp


@pytest.mark.parametrize('name', CASES)
def test_parameters_match_reference(name):
    g = load_golden(name)
    wf = build(g)
    assert [k for k, _ in wf.named_parameters()] == list(g['param_names'])
    assert wf.param_num == int(g['param_num'])
    sums = np.array([[float(p.sum()), float((p * p).sum())] for p in wf.parameters()])
    assert np.allclose(sums, g['init_checksums'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    wf = build(g)
    Wa, ba, Wp, bp = numpy_weights(wf)
    masks = onp.NumberSpinMasks(int(g['qubit_num']), int(g['particle_num']))
    x = g['samples'].view(np.uint64)
    nphys = int(g['n_phys'])
    lp = onp.nade_log_psi(x, masks, Wa, ba, Wp, bp)
    assert np.abs(lp[:nphys] - g['log_psi'][:nphys]).max() < 1e-12
    assert np.abs(np.exp(lp) - g['amplitude']).max() < 1e-12
    for key in [k for k in g if k.startswith('cond_log_abs_q')]:
        q = int(key.split('q')[-1])
        c = onp.nade_cond_log_abs(x[:nphys], q, masks, Wa[q], ba[q])
        ref = g[key]
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < 1e-12
