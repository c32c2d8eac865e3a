"""GPU parity, NADE mode (the reference's default de_mode): nade_forward_kernel through LogAbsPhaseANQS(de_mode='NADE')
against the golden vectors of the unmodified reference.  1e-10 on log psi / conditionals / gradients; sampled
configurations and counts bit-exact given identical draws."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_oracle_nade import build, CASES

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('name', CASES)
def test_log_psi_cond_and_gradients(name):
    g = load_golden(name)
    wf = build(g, device=DEV)
    nphys = int(g['n_phys'])
    s = _dev(g['samples']).view(-1, 1)
    with torch.no_grad():
        lp = wf.log_psi_of_indices(s).cpu().numpy()
        amp = wf.amplitude(s).cpu().numpy()
    assert np.abs(lp[:nphys] - g['log_psi'][:nphys]).max() < 1e-10
    assert np.array_equal(np.isneginf(lp.real), np.isneginf(g['log_psi'].real))
    assert np.abs(amp - g['amplitude']).max() < 1e-10
    for key in [k for k in g if k.startswith('cond_log_abs_q')]:
        q = int(key.split('q')[-1])
        c = wf.cond_log_abs(qudit_idx=q, prefix_idx=s[:nphys].view(-1))[:, :g[key].shape[1]].cpu().numpy()
        ref = g[key]
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < 1e-10
    # gradients against the reference's autograd
    c = _dev(g['grad_coeff'])
    proj = np.random.default_rng(int(g['seed']) + 3).standard_normal((16, wf.param_num))
    wf.zero_grad()
    loss = (torch.conj(c) * wf.log_psi_of_indices(s[:nphys])).real.sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(g['grad_loss'])) < 1e-9 * max(1.0, abs(float(g['grad_loss'])))
    grad = wf.cat_grad.cpu().numpy()
    assert np.abs(proj @ grad - g['grad_proj']).max() < 1e-10 * max(1.0, np.abs(g['grad_proj']).max())
    norms = np.array([float(p.grad.norm()) for p in wf.parameters()])
    assert np.abs(norms - g['grad_norms']).max() < 1e-10 * max(1.0, g['grad_norms'].max())


@pytest.mark.parametrize('name', CASES)
def test_samplers_match_reference(name):
    g = load_golden(name)
    wf = build(g, device=DEV)
    idx, cnt = wf.sample_stats(int(g['stats_num']), draw_mode='rint')
    assert np.array_equal(idx.view(-1).cpu().numpy(), g['stats_idx'])
    assert np.array_equal(cnt.real.cpu().numpy(), g['stats_counts'])
    urng = np.random.default_rng(int(g['seed']) + 4)
    gi, gf = wf.sample_indices_gumbel(int(g['gumbel_num']), uniforms=lambda q, B, D: torch.from_numpy(urng.random((B, D))))
    assert np.array_equal(gi.view(-1).cpu().numpy(), g['gumbel_idx'])
    assert np.abs(gf.cpu().numpy() - g['gumbel_freqs']).max() < 1e-10
    # Philox mode: conservation and physicality
    idx, cnt = wf.sample_stats(10 ** 5, seed=9)
    assert float(cnt.real.sum()) == 1e5 and float(cnt.real.min()) >= 1.0


def test_log_jacobian_matches_autograd():
    """compute_cat_log_jac (batched outer products) against one autograd pass per sample, NADE mode."""
    g = load_golden('nade_n12')
    wf = build(g, device=DEV)
    s = _dev(g['samples'][:5]).view(-1, 1)
    jac = wf.compute_cat_log_jac(s)
    params = list(wf.parameters())
    for b in range(5):
        lp = wf.log_psi_of_indices(s[b:b + 1])
        g_re = torch.autograd.grad(lp.real.sum(), params, retain_graph=True)
        g_im = torch.autograd.grad(lp.imag.sum(), params)
        row = torch.complex(torch.cat([x.reshape(-1) for x in g_re]), -torch.cat([x.reshape(-1) for x in g_im]))
        assert (jac[b] - row).abs().max() < 1e-12


@pytest.mark.parametrize('name', CASES)
def test_tensor_core_mode_within_tolerance(name):
    """NADE on the tensor cores (tcgen05 tf32, k3_nade_tc.cu) against the fp64 kernel: identical masks, log|psi| within 2e-2,
    phase within 1e-1 rad, conditionals within 2e-2; sampling through it conserves counts and stays physical; a parameter update
    is picked up."""
    g = load_golden(name)
    wf = build(g, device=DEV)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    s = _dev(g['samples']).view(-1, 1)
    unphys = torch.full((3, 1), (1 << n) - 1, dtype=torch.int64, device=DEV)   # all orbitals occupied: outside the sector
    idx = torch.cat((s, unphys))
    with torch.no_grad():
        ref = wf.log_psi_of_indices(idx)
        wf.set_inference_precision('tf32')
        tc = wf.log_psi_of_indices(idx)
    fin = torch.isfinite(ref.real)
    assert torch.equal(fin, torch.isfinite(tc.real)) and 0 < int(fin.sum()) <= s.shape[0]   # the golden samples hold unphysical ones too
    assert float((tc.real[fin] - ref.real[fin]).abs().max()) < 2e-2
    assert float((tc.imag[fin] - ref.imag[fin]).abs().max()) < 1e-1
    for q in range(wf.qudit_num):
        wf.set_inference_precision('fp64')
        c0 = wf.cond_log_abs(qudit_idx=q, prefix_idx=s.view(-1))
        wf.set_inference_precision('tf32')
        c1 = wf.cond_log_abs(qudit_idx=q, prefix_idx=s.view(-1))
        f = torch.isfinite(c0)
        assert torch.equal(f, torch.isfinite(c1))
        assert float((c0[f] - c1[f]).abs().max()) < 2e-2
    si, sc = wf.sample_stats(10 ** 5, seed=3)
    assert float(sc.real.sum()) == 1e5
    even = torch.tensor(0x5555555555555555, dtype=torch.int64, device=DEV)
    assert bool((wf.hilbert_space.popcount(si.view(-1) & even) == ne // 2).all())
    with torch.no_grad():
        for p in wf.parameters():
            p.add_(0.01 * torch.randn_like(p))
        tc2 = wf.log_psi_of_indices(s)
        wf.set_inference_precision('fp64')
        ref2 = wf.log_psi_of_indices(s)
    f2 = torch.isfinite(ref2.real)
    assert float((tc2.real[f2] - ref2.real[f2]).abs().max()) < 2e-2 and float((ref2.real[f2] - ref.real[:s.shape[0]][f2]).abs().max()) > 1e-3
    # gradients never take the tensor-core path
    wf.set_inference_precision('tf32')
    lp = wf.log_psi_of_indices(s[:4])
    assert lp.requires_grad and torch.equal(lp.detach(), ref2[:4])
