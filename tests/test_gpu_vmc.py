"""GPU parity: one whole VMC iteration (sample -> amplitudes -> local energies -> loss -> backward) through the drop-in
objects and the mirrored call sites of `calculations`, against the same iteration run by the unmodified reference
(tests/golden/vmc_*.npz, made by oracle/make_golden.py: group 'vmc').  Sampled configurations and counts bit-exact,
energies / loss / gradient within 1e-10 (relative to their scale)."""
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry,
                                         SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig,
                                         SamplingConfig, SamplingResult, sample, LocalEnergyCalculationConfig,
                                         compute_local_energies, vmc_loss, synthetic, SRConfig, ProcessGradConfig, process_grad)
from oracle.make_golden import made_weights

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def build(g):
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=int(g['n_irreps']), seed=int(g['ham_seed']))
    tmp = tempfile.mkdtemp(prefix='anqs_vmc_test_')
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
    nets = made_weights(n, wf.qudit_num, wf.max_qudit_dim, seed=int(g['weight_seed']))
    sd = {}
    for name, layers in zip(('log_abs_subnet', 'phase_subnet'), nets):
        for l, (wt, b) in enumerate(layers):
            sd[f'{name}.layers.{l}.weight'] = torch.from_numpy(wt.copy())
            sd[f'{name}.layers.{l}.bias'] = torch.from_numpy(b.copy())
    wf.load_state_dict(sd)
    return hs, ham, wf


@pytest.mark.parametrize('name', ['vmc_n12', 'vmc_n20'])
def test_vmc_iteration_matches_reference(name):
    g = load_golden(name)
    hs, ham, wf = build(g)
    # EXP:504-522: sample, then sort ascending (unsigned)
    res, unq_num, reps, _ = sample(wf=wf, config=SamplingConfig(sample_indices=False, sample_num=int(g['sample_num'])), draw_mode='rint')
    indices, perm = wf.sort_base_idx(res.indices)
    counts = res.counts[perm]
    assert np.array_equal(indices.view(-1).cpu().numpy(), g['indices'])
    assert np.array_equal(counts.real.cpu().numpy(), g['counts'])
    # EXP:548-611 with loss_type='sample_aware_e_loc'
    wf.zero_grad()
    amps = wf.amplitude(indices)
    assert np.abs(amps.detach().cpu().numpy() - g['amps']).max() < 1e-12
    sr = SamplingResult(indices=indices, counts=counts)
    le, _ = compute_local_energies(wf=wf, sampling_result=sr, sampled_amps=amps.detach(), ham=ham,
                                   config=LocalEnergyCalculationConfig(use_tree_for_candidates='ham'), sample_aware=True)
    est = le.sample_aware_e_loc_mc_est
    scale = max(1.0, np.abs(g['eloc']).max())
    assert np.abs(est.values.cpu().numpy() - g['eloc']).max() < 1e-10 * scale
    assert abs(complex(est.mean) - complex(g['energy_mean'])) < 1e-10 * scale
    assert abs(complex(est.var) - complex(g['energy_var'])) < 1e-10 * scale * scale
    loss = vmc_loss(amps, est)
    assert abs(float(loss.detach()) - float(g['loss'])) < 1e-10 * max(1.0, abs(float(g['loss'])))
    loss.backward()
    grad = wf.cat_grad.cpu().numpy()
    proj = np.random.default_rng(int(g['weight_seed']) + 3).standard_normal((16, grad.shape[0]))
    gs = max(1.0, float(g['grad_norm']))
    assert abs(np.linalg.norm(grad) - float(g['grad_norm'])) < 1e-10 * gs
    assert np.abs(proj @ grad - g['grad_proj']).max() < 1e-9 * gs
    assert np.abs(grad[:64] - g['grad_head']).max() < 1e-10 * gs and np.abs(grad[-64:] - g['grad_tail']).max() < 1e-10 * gs
    # gradient post-processing (PG:55-70): per-sample log-Jacobian, SR on the top-25 samples, clipping
    jac = wf.compute_cat_log_jac(indices[:8]).cpu().numpy()
    jproj = np.random.default_rng(int(g['weight_seed']) + 5).standard_normal((grad.shape[0], 4))
    assert np.abs(jac[:, :32] - g['log_jac_head']).max() < 1e-10
    assert np.abs(jac @ jproj - g['log_jac_proj']).max() < 1e-9 * max(1.0, np.abs(g['log_jac_proj']).max())
    raw = wf.cat_grad.clone()
    process_grad(wf=wf, sampling_result=sr, config=ProcessGradConfig())
    g_sr = wf.cat_grad.cpu().numpy()
    assert abs(np.linalg.norm(g_sr) - float(g['sr_grad_norm'])) < 1e-8
    assert np.abs(proj @ g_sr - g['sr_grad_proj']).max() < 1e-7 * max(1.0, np.abs(g['sr_grad_proj']).max())
    assert np.abs(g_sr[:64] - g['sr_grad_head']).max() < 1e-7
    wf.cat_grad = raw
    process_grad(wf=wf, sampling_result=sr, config=ProcessGradConfig(sr_config=SRConfig(use_reg=False, max_indices_num=10),
                                                                      clip_grad_norm=False, renorm_grad=True))
    g_sr2 = wf.cat_grad.cpu().numpy()
    assert abs(np.linalg.norm(g_sr2) - 1.0) < 1e-12
    assert np.abs(proj @ g_sr2 - g['sr2_grad_proj']).max() < 1e-6 * max(1.0, np.abs(g['sr2_grad_proj']).max())
    # the full (not sample-aware) local energy: de-duplicated non-sampled x', amplitudes from the network (PO:992-1105)
    for version in ('old', 'new'):
        le_full, metrics = compute_local_energies(wf=wf, sampling_result=sr, sampled_amps=amps.detach(), ham=ham,
                                                  config=LocalEnergyCalculationConfig(use_tree_for_candidates='ham', code_version=version),
                                                  sample_aware=False)
        assert np.abs(le_full.full_e_loc_mc_est.values.cpu().numpy() - g['eloc_full']).max() < 1e-10 * scale
        assert np.abs(le_full.sample_aware_e_loc_mc_est.values.cpu().numpy() - g['eloc_full_aware']).max() < 1e-10 * scale
        assert abs(complex(le_full.full_e_loc_mc_est.mean) - complex(g['full_energy_mean'])) < 1e-10 * scale
        assert metrics.non_sampled_unq_x_primes_num == int(g['non_sampled_unq'])
        assert metrics.candidate_x_primes_num == int(g['candidates'])


def test_exact_energy_when_the_sector_is_exhausted():
    """BASELINE config 2 (H2O STO-3G shape, 14 qubits, 10 electrons): 1e5 samples visit all 441 states of the sector, so
    the sample-aware energy equals <psi|H|psi> / <psi|psi> on the dense sector matrix, and it is bounded below by the
    lowest sector eigenvalue (the stand-in for FCI)."""
    n, ne = 14, 10
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
    tmp = tempfile.mkdtemp(prefix='anqs_vmc_test_')
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(0)
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
    res, _, _, _ = sample(wf=wf, config=SamplingConfig(sample_indices=False, sample_num=10 ** 5), seed=11)
    indices, perm = wf.sort_base_idx(res.indices)
    assert indices.shape[0] == 441 and float(res.counts.real.sum()) == 1e5
    with torch.no_grad():
        amps = wf.amplitude(indices)
    le, _ = compute_local_energies(wf=wf, sampling_result=SamplingResult(indices=indices, counts=res.counts[perm]), sampled_amps=amps,
                                   ham=ham, config=LocalEnergyCalculationConfig(use_tree_for_candidates='ham'), sample_aware=True)
    # dense sector matrix straight from the Pauli arrays: H[x, x'] = sum_t w_t (-1)^popcount(x' & yz_t), x = x' ^ xy_t
    idx = indices.view(-1).cpu().numpy().view(np.uint64)
    Hs = np.zeros((441, 441), dtype=np.complex128)
    cols = np.arange(441)
    for xy_t, yz_t, w_t in zip(xy.view(np.uint64), yz.view(np.uint64), w):
        x = idx ^ xy_t
        pos = np.searchsorted(idx, x)
        ok = (pos < 441) & (idx[np.minimum(pos, 440)] == x)
        par = np.array([bin(int(v)).count('1') & 1 for v in (idx & yz_t)])
        np.add.at(Hs, (pos[ok], cols[ok]), w_t * (1.0 - 2.0 * par[ok]))
    psi = amps.cpu().numpy()
    e_exact = (np.conj(psi) @ (Hs @ psi)) / (np.conj(psi) @ psi)
    e = complex(le.sample_aware_e_loc_mc_est.mean)
    assert abs(e - e_exact) < 1e-10 * max(1.0, abs(e_exact))
    assert np.abs(Hs - Hs.conj().T).max() < 1e-12
    assert e.real >= np.linalg.eigvalsh(Hs)[0] - 1e-10


@pytest.mark.parametrize('n', [0, 1, 31, 1000, 1_000_003])
def test_energy_stats_kernel_matches_the_expression(n):
    """dist.local_energy_stats on device tensors (energy_stats_kernel, one pass) against the elementwise expression it
    replaces (MonteCarloEstimator sums, CLE:48-62): 1e-12 relative, the same bits on a second call."""
    from anqs_quantum_chemistry_b200 import dist as adist
    g = torch.Generator().manual_seed(n + 1)
    e = torch.complex(torch.randn(n, generator=g, dtype=torch.float64) * 50 - 100, torch.randn(n, generator=g, dtype=torch.float64))
    a = torch.complex(torch.randn(n, generator=g, dtype=torch.float64), torch.randn(n, generator=g, dtype=torch.float64)) * 1e-3
    ref = adist.local_energy_stats(e, a)                      # CPU tensors: the expression
    out = adist.local_energy_stats(e.to(DEV), a.to(DEV))
    out2 = adist.local_energy_stats(e.to(DEV), a.to(DEV))
    assert out.is_cuda and torch.equal(out, out2)
    scale = float(ref.abs().max()) if n else 1.0
    assert float((out.cpu() - ref).abs().max()) <= 1e-12 * max(scale, 1e-300)
    if n:
        mean, var, norm = adist.reduce_energy_stats(out, world_size=1)
        w = (a.abs() ** 2).double()
        m_ref = (w * e).sum() / w.sum()
        assert abs(complex(mean.cpu()) - complex(m_ref)) < 1e-10 * abs(complex(m_ref))
