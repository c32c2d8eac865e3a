"""GPU parity: the MADE wave-function kernels (k3) and the sampler kernels (k4), driven through the LogAbsPhaseANQS
drop-in and the C ABI, against the golden vectors written by the unmodified reference (tests/golden/anqs_*.npz) and
against the numpy oracle on larger seeded inputs.  Tolerances: 1e-10 on log psi / amplitudes / gradients (fp64);
sampled configurations and counts bit-exact given identical draws."""
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, Z2Symmetry,
                                         LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, LocalSamplingConfig, synthetic)
from oracle import anqs_numpy as onp
from oracle.make_golden import made_weights

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
CASES = ['anqs_n12', 'anqs_n14', 'anqs_n20', 'anqs_n56']
# the reference's default symmetry level (Z2 generators on top of N and S_z, create_masker.py:18-24) and a non-zero
# masking depth (the last qudit sampled unmasked, ANQS:417-418, 605-606) reach the kernels through the same descriptors
CASES_SYM = CASES + ['anqs_z2_n12', 'anqs_md1_n20']


def build(n, ne, nets=None, seed=0, z2=(), masking_depth=0):
    tmp = tempfile.mkdtemp(prefix='anqs_gpu_test_')
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=tmp, rng_seed=seed)
    syms = (ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne), SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0))
    syms += tuple(Z2Symmetry(hilbert_space=hs, value=int(v), pauli_z_positions=[i for i in range(n) if (int(m) >> i) & 1]) for v, m in z2)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=syms)
    torch.manual_seed(seed)
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker,
                         config=ANQSConfig(de_mode='MADE', local_sampling_config=LocalSamplingConfig(masking_depth=masking_depth)))
    if nets is not None:
        sd = {}
        for name, layers in zip(('log_abs_subnet', 'phase_subnet'), nets):
            for l, (w, b) in enumerate(layers):
                sd[f'{name}.layers.{l}.weight'] = torch.from_numpy(w.copy())
                sd[f'{name}.layers.{l}.bias'] = torch.from_numpy(b.copy())
        wf.load_state_dict(sd)
    return hs, masker, wf


def setup_case(name):
    g = load_golden(name)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    masks = onp.NumberSpinMasks(n, ne)
    nets = made_weights(n, int(g['qudit_num']), int(g['max_qudit_dim']), seed=int(g['weight_seed']))
    z2 = tuple(zip(g['z2_values'].tolist(), g['z2_masks'].tolist())) if 'z2_values' in g else ()
    hs, masker, wf = build(n, ne, nets, z2=z2, masking_depth=int(g['masking_depth']) if 'masking_depth' in g else 0)
    return g, masks, nets, wf


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('name', CASES_SYM)
def test_log_psi_and_amplitude_match_reference(name):
    g, masks, nets, wf = setup_case(name)
    s = _dev(g['samples']).view(-1, 1)
    nphys = int(g['n_phys'])
    with torch.no_grad():
        lp = wf.log_psi_of_indices(s).cpu().numpy()
        amp = wf.amplitude(s).cpu().numpy()
        lp_vec = wf.log_psi(wf.base_idx2base_vec(s)).cpu().numpy()
    assert np.abs(lp[:nphys] - g['log_psi'][:nphys]).max() < 1e-10
    # unphysical configurations: log|psi| = -inf, amplitude exactly 0 (ANQS:399-401)
    assert np.array_equal(np.isneginf(lp.real), np.isneginf(g['log_psi'].real))
    assert np.abs(amp - g['amplitude']).max() < 1e-10
    assert np.array_equal(lp_vec, lp)


@pytest.mark.parametrize('name', CASES_SYM)
def test_cond_log_abs_matches_reference(name):
    g, masks, nets, wf = setup_case(name)
    nphys = int(g['n_phys'])
    s = _dev(g['samples'][:nphys])
    for key in [k for k in g if k.startswith('cond_log_abs_q')]:
        q = int(key.split('q')[-1])
        c = wf.cond_log_abs(qudit_idx=q, prefix_idx=s).cpu().numpy()   # bits above the prefix are ignored by the kernel
        ref = g[key]
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < 1e-10
        # the reference signature (bit-vector prefix) gives the same numbers
        vec = wf.base_idx2base_vec(s.view(-1, 1))[:, :wf.qudit_starts[q]]
        c2 = wf.cond_log_abs(qudit_idx=q, base_vec=vec).cpu().numpy()
        assert np.array_equal(c2, c)


@pytest.mark.parametrize('name', CASES_SYM)
def test_gradients_match_reference_autograd(name):
    g, masks, nets, wf = setup_case(name)
    nphys = int(g['n_phys'])
    s = _dev(g['samples'][:nphys]).view(-1, 1)
    c = _dev(g['grad_coeff'])
    proj = np.random.default_rng(int(g['weight_seed']) + 3).standard_normal((16, wf.param_num))
    wf.zero_grad()
    lp = wf.log_psi_of_indices(s)
    loss = (torch.conj(c) * lp).real.sum()
    loss.backward()
    grad = wf.cat_grad.cpu().numpy()
    assert abs(float(loss) - float(g['grad_loss'])) < 1e-9 * max(1.0, abs(float(g['grad_loss'])))
    scale = max(1.0, np.abs(g['grad_proj']).max())
    assert np.abs(proj @ grad - g['grad_proj']).max() < 1e-10 * scale
    norms = np.array([float(p.grad.norm()) for p in wf.parameters()])
    assert np.abs(norms - g['grad_norms']).max() < 1e-10 * max(1.0, g['grad_norms'].max())   # includes Q4: grads on masked weights
    assert np.abs(grad[:64] - g['grad_head']).max() < 1e-10 * scale
    assert np.abs(grad[-64:] - g['grad_tail']).max() < 1e-10 * scale
    # amplitude path (exp on top of log psi, ANQS:483-485)
    wf.zero_grad()
    loss2 = (torch.conj(c) * wf.amplitude(s)).real.sum()
    loss2.backward()
    ga = wf.cat_grad.cpu().numpy()
    assert np.abs(proj @ ga - g['grad_amp_proj']).max() < 1e-10 * max(1.0, np.abs(g['grad_amp_proj']).max())


@pytest.mark.parametrize('name', CASES_SYM)
def test_sample_stats_rint_matches_reference(name):
    g, masks, nets, wf = setup_case(name)
    md = int(g['masking_depth']) if 'masking_depth' in g else 0
    for call in range(2):   # the second call sizes its levels from the first one's (one host read): same result
        idx, cnt = wf.sample_stats(int(g['stats_num']), draw_mode='rint')
        assert idx.dtype == torch.int64 and idx.dim() == 2 and cnt.dtype == torch.complex128
        assert np.array_equal(idx.view(-1).cpu().numpy(), g['stats_idx'])
        assert np.array_equal(cnt.real.cpu().numpy(), g['stats_counts'])
        if md == 0:
            assert float(cnt.real.sum()) == float(g['stats_num'])
        else:   # an unmasked level ('DU') loses the samples that fell on unphysical children (ANQS:653-655)
            assert float(cnt.real.sum()) < float(g['stats_num'])


@pytest.mark.parametrize('name', CASES_SYM)
def test_gumbel_matches_reference(name):
    g, masks, nets, wf = setup_case(name)
    urng = np.random.default_rng(int(g['weight_seed']) + 4)
    idx, freqs = wf.sample_indices_gumbel(int(g['gumbel_num']), uniforms=lambda q, B, D: torch.from_numpy(urng.random((B, D))))
    assert np.array_equal(idx.view(-1).cpu().numpy(), g['gumbel_idx'])
    assert np.abs(freqs.cpu().numpy() - g['gumbel_freqs']).max() < 1e-10


@pytest.mark.parametrize('n,ne', [(12, 4), (14, 10)])
def test_normalisation_over_physical_sector(n, ne):
    """sum over the (N, S_z) sector of |psi|^2 is 1; everything outside has amplitude 0."""
    hs, masker, wf = build(n, ne)
    allx = torch.arange(2 ** n, dtype=torch.int64, device=DEV).view(-1, 1)
    with torch.no_grad():
        amp = wf.amplitude(allx)
    p = (amp.abs() ** 2).cpu().numpy()
    x = np.arange(2 ** n, dtype=np.uint64)
    ev = np.array([bin(int(v) & 0x5555555555555555).count('1') for v in x])
    od = np.array([bin(int(v) & 0xAAAAAAAAAAAAAAAA).count('1') for v in x])
    phys = (ev == ne // 2) & (od == ne // 2)
    assert abs(p[phys].sum() - 1.0) < 1e-12
    assert p[~phys].max() == 0.0


@pytest.mark.parametrize('name', ['anqs_n20', 'anqs_n56'])
def test_log_psi_matches_oracle_on_large_batch(name):
    g, masks, nets, wf = setup_case(name)
    n, ne = masks.n, masks.particle_num
    x = synthetic.random_physical_samples(n, ne // 2, ne // 2, 5000 if n == 20 else 20011, seed=11)
    Wa, ba, Wp, bp = onp.masked_weights(nets, masks)
    ref = onp.log_psi(x, masks, Wa, ba, Wp, bp)
    with torch.no_grad():
        lp = wf.log_psi_of_indices(_dev(x.view(np.int64))).cpu().numpy()
    assert np.abs(lp - ref).max() < 1e-10


def test_sample_stats_philox_properties():
    """Count conservation, physicality, reproducibility, distribution."""
    hs, masker, wf = build(20, 14)
    N = 10 ** 7
    idx, cnt = wf.sample_stats(N, seed=123)
    idx2, cnt2 = wf.sample_stats(N, seed=123)
    assert torch.equal(idx, idx2) and torch.equal(cnt, cnt2)
    idx3, cnt3 = wf.sample_stats(N, seed=124)
    assert not (idx3.shape == idx.shape and torch.equal(cnt3, cnt))
    c = cnt.real
    assert float(c.sum()) == float(N) and float(c.min()) >= 1.0 and torch.equal(c, c.round())
    x = idx.view(-1)
    assert torch.equal(x.unique(), x.sort().values)                       # unique configurations
    ev = hs.popcount(x & 0x5555555555555555)
    od = hs.popcount(x & ~0x5555555555555555)
    assert bool((ev == 7).all()) and bool((od == 7).all())
    # empirical frequencies follow |psi|^2: chi-square per degree of freedom close to 1
    with torch.no_grad():
        p = wf.amplitude(idx).abs() ** 2
    expect = p * N
    big = expect > 20
    chi2 = float((((c - expect) ** 2) / expect)[big].sum() / big.sum())
    assert 0.8 < chi2 < 1.2, chi2


def test_unmasked_level_sampling_philox():
    """LocalSamplingConfig(masking_depth=1): the last qudit is drawn from its UNMASKED conditionals and the unphysical children
    are dropped with their samples (ANQS:605-606, 653-655; Gumbel: ANQS:708-709, 804-809).  Production (Philox) draws: only
    physical, unique configurations come back, fewer samples than asked for, reproducibly, and the counts follow the product
    of the conditionals the sampler drew from."""
    hs, masker, wf = build(20, 14, masking_depth=1)
    qg = wf.qubit_grouping
    N = 4 * 10 ** 6
    idx, cnt = wf.sample_stats(N, seed=11)
    idx2, cnt2 = wf.sample_stats(N, seed=11)           # second call: predicted level sizes, one host read
    assert torch.equal(idx, idx2) and torch.equal(cnt, cnt2)
    c, x = cnt.real, idx.view(-1)
    assert 0 < float(c.sum()) < float(N) and float(c.min()) >= 1.0 and torch.equal(c, c.round())
    assert x.unique().shape[0] == x.shape[0]
    assert bool((hs.popcount(x & 0x5555555555555555) == 7).all()) and bool((hs.popcount(x & ~0x5555555555555555) == 7).all())
    logp = torch.zeros(x.shape[0], dtype=torch.float64, device=DEV)
    for q in range(qg.qudit_num):
        start, D = qg.qudit_starts[q], qg.qudit_dims_host[q]
        cond = wf.cond_log_abs(qudit_idx=q, prefix_idx=x & ((1 << start) - 1))
        logp += 2.0 * cond.gather(1, ((x >> start) & (D - 1)).view(-1, 1)).view(-1)
    expect = torch.exp(logp) * N
    assert float(expect.sum()) < N                     # probability mass sits on unphysical children of the unmasked level
    big = expect > 20
    chi2 = float((((c - expect) ** 2) / expect)[big].sum() / big.sum())
    assert 0.8 < chi2 < 1.2, chi2
    # Gumbel top-k: dead rows carried to the end (default) and compacted after every level give the same set and frequencies
    gi, gf = wf.sample_indices_gumbel(3000, seed=5)
    gi2, gf2 = wf.sample_indices_gumbel(3000, seed=5, compact_levels=True)
    g = gi.view(-1)
    assert 0 < g.shape[0] <= 3000 and g.unique().shape[0] == g.shape[0]
    assert bool((hs.popcount(g & 0x5555555555555555) == 7).all()) and bool((hs.popcount(g & ~0x5555555555555555) == 7).all())
    o1, o2 = torch.argsort(g), torch.argsort(gi2.view(-1))
    assert torch.equal(g[o1], gi2.view(-1)[o2]) and torch.allclose(gf[o1], gf2[o2], rtol=0, atol=1e-12)
    assert abs(float(gf.sum()) - 1.0) < 1e-12


def test_sample_stats_philox_small_counts():
    """Inversion branch of the binomial generator (n p < 10) and the 56-qubit tree (10 levels)."""
    hs, masker, wf = build(56, 14)
    idx, cnt = wf.sample_stats(5000, seed=7)
    c = cnt.real
    assert float(c.sum()) == 5000.0 and float(c.min()) >= 1.0
    x = idx.view(-1)
    assert bool((hs.popcount(x & 0x5555555555555555) == 7).all()) and bool((hs.popcount(x & ~0x5555555555555555) == 7).all())
    assert x.unique().shape[0] == x.shape[0]


def test_gumbel_philox_properties():
    hs, masker, wf = build(20, 14)
    idx, freqs = wf.sample_indices_gumbel(2000, seed=5)
    x = idx.view(-1)
    assert x.shape[0] == 2000 and x.unique().shape[0] == 2000
    assert bool((hs.popcount(x & 0x5555555555555555) == 7).all()) and bool((hs.popcount(x & ~0x5555555555555555) == 7).all())
    assert abs(float(freqs.sum()) - 1.0) < 1e-12
    with torch.no_grad():
        p = wf.amplitude(idx).abs() ** 2
    assert torch.allclose(freqs, p / p.sum(), rtol=0, atol=1e-12)
    idx2, _ = wf.sample_indices_gumbel(2000, seed=5)
    assert torch.equal(idx, idx2)


def test_cat_log_jac_matches_finite_differences():
    hs, masker, wf = build(12, 4)
    x = _dev(synthetic.random_physical_samples(12, 2, 2, 6, seed=3).view(np.int64)).view(-1, 1)
    jac = wf.compute_cat_log_jac(x)
    assert tuple(jac.shape) == (6, wf.param_num) and jac.dtype == torch.complex128
    params = list(wf.parameters())
    # central differences on a few unmasked first-layer biases and last-layer weights
    for p, flat_off in ((params[1], sum(q.numel() for q in params[:1])), (params[5], sum(q.numel() for q in params[:5]))):
        for j in (0, p.numel() - 1):
            eps = 1e-6
            with torch.no_grad():
                old = p.view(-1)[j].item()
                p.view(-1)[j] = old + eps
                lp_p = wf.log_psi_of_indices(x)
                p.view(-1)[j] = old - eps
                lp_m = wf.log_psi_of_indices(x)
                p.view(-1)[j] = old
            fd = torch.conj((lp_p - lp_m) / (2 * eps))
            assert torch.allclose(jac[:, flat_off + j], fd, rtol=0, atol=1e-7)


def test_empty_and_ragged_batches():
    hs, masker, wf = build(12, 4)
    with torch.no_grad():
        assert wf.amplitude(torch.zeros((0, 1), dtype=torch.int64, device=DEV)).shape[0] == 0
        x = _dev(synthetic.random_physical_samples(12, 2, 2, 65, seed=4).view(np.int64)).view(-1, 1)
        a65 = wf.amplitude(x)
        a1 = torch.cat([wf.amplitude(x[i:i + 1]) for i in range(65)])
    assert torch.equal(a65, a1)   # tile boundaries (64 samples per tile) do not change results


# ---- tensor-core (tcgen05, tf32) inference mode ---------------------------------------------------------------------
TC_TOL_LOG_ABS, TC_TOL_PHASE = 2e-2, 1e-1   # absolute, on log|psi| and arg psi (rad); tf32 products carry 10 mantissa bits


@pytest.mark.parametrize('name', CASES)
def test_tensor_core_log_psi_within_tolerance(name):
    g, masks, nets, wf = setup_case(name)
    s = _dev(g['samples']).view(-1, 1)
    nphys = int(g['n_phys'])
    lp = wf.log_psi_tc(s).cpu().numpy()
    ref = g['log_psi']
    assert np.array_equal(np.isneginf(lp.real), np.isneginf(ref.real))          # unphysical configurations exactly
    assert np.abs(lp.real[:nphys] - ref.real[:nphys]).max() < TC_TOL_LOG_ABS
    assert np.abs(lp.imag[:nphys] - ref.imag[:nphys]).max() < TC_TOL_PHASE
    # typical error is far below the bound
    assert np.abs(lp.real[:nphys] - ref.real[:nphys]).mean() < 5e-3
    # routing: no-grad amplitudes use the tensor cores once the mode is switched on; gradients stay fp64
    wf.set_inference_precision('tf32')
    with torch.no_grad():
        amp = wf.amplitude(s).cpu().numpy()
    assert np.abs(amp - np.exp(lp)).max() < 1e-14
    lp64 = wf.log_psi_of_indices(s[:nphys])
    assert lp64.requires_grad and np.abs(lp64.detach().cpu().numpy() - ref[:nphys]).max() < 1e-10


@pytest.mark.parametrize('name', ['anqs_n12', 'anqs_n56'])
def test_tensor_core_cond_log_abs_and_sampling(name):
    g, masks, nets, wf = setup_case(name)
    nphys = int(g['n_phys'])
    s = _dev(g['samples'][:nphys])
    wf.set_inference_precision('tf32')
    for key in [k for k in g if k.startswith('cond_log_abs_q')]:
        q = int(key.split('q')[-1])
        c = wf.cond_log_abs(qudit_idx=q, prefix_idx=s).cpu().numpy()
        ref = g[key]
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))                 # masks are exact
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < TC_TOL_LOG_ABS
        assert np.abs(np.exp(2 * c).sum(axis=1) - 1.0).max() < 1e-5              # conditionals stay normalised
    idx, cnt = wf.sample_stats(10 ** 5, seed=3)
    c = cnt.real
    assert float(c.sum()) == 1e5 and float(c.min()) >= 1.0
    x = idx.view(-1)
    ne = int(g['particle_num']) // 2
    hs = wf.hilbert_space
    assert bool((hs.popcount(x & 0x5555555555555555) == ne).all()) and bool((hs.popcount(x & ~0x5555555555555555) == ne).all())


def test_tensor_core_batch_shapes_and_repack():
    """Tile boundaries (128 samples per CTA tile), empty batch, and repacking after a parameter update."""
    hs, masker, wf = build(20, 14)
    x = _dev(synthetic.random_physical_samples(20, 7, 7, 1000, seed=4).view(np.int64)).view(-1, 1)
    assert wf.log_psi_tc(x[:0]).shape[0] == 0
    full = wf.log_psi_tc(x)
    for b in (1, 127, 128, 129, 333):
        assert torch.equal(wf.log_psi_tc(x[:b]), full[:b])
    with torch.no_grad():
        ref = wf.log_psi_of_indices(x)
    assert (full.real - ref.real).abs().max() < TC_TOL_LOG_ABS and (full.imag - ref.imag).abs().max() < TC_TOL_PHASE
    with torch.no_grad():
        for p in wf.parameters():
            p.add_(0.05 * torch.randn_like(p))
        ref2 = wf.log_psi_of_indices(x)
    full2 = wf.log_psi_tc(x)
    assert (full2.real - ref2.real).abs().max() < TC_TOL_LOG_ABS and (full2.imag - ref2.imag).abs().max() < TC_TOL_PHASE
    assert (full2 - full).abs().max() > 1e-3


@pytest.mark.parametrize('n,ne,samples', [(20, 14, 10 ** 6), (56, 14, 20000)])
def test_sub_tree_sharded_sampling_is_identical(n, ne, samples):
    """the count-splitting tree sharded by sub-tree over 1, 2, 3 and 8 (emulated) ranks gives,
    concatenated in rank order, exactly the single-GPU result - the draws are keyed by the node's prefix."""
    from anqs_quantum_chemistry_b200 import dist as adist
    hs, masker, wf = build(n, ne)
    ref_idx, ref_cnt = wf.sample_stats(samples, seed=77)
    for world in (1, 2, 3, 8):
        parts = [adist.sharded_sample_stats(wf, samples, seed=77, world_size=world, rank=r, gather=False, min_nodes_per_rank=4)
                 for r in range(world)]
        idx = torch.cat([p[0] for p in parts])
        cnt = torch.cat([p[1] for p in parts])
        assert torch.equal(idx, ref_idx) and torch.equal(cnt, ref_cnt), world
        sizes = [p[0].shape[0] for p in parts]
        assert min(sizes) > 0


@pytest.mark.parametrize('n,ne,num', [(12, 4, 100), (20, 14, 5000), (14, 10, 10 ** 4)])
def test_gumbel_dead_rows_equal_compaction(n, ne, num):
    """Production Gumbel sampling carries masked children on as dead rows, takes an UNSORTED top-k per level and reads the host
    once; compacting and sorting after every level (what the reference does, and what the parity mode with injected uniforms
    does) gives the same set of samples with the same frequencies - the draws are keyed by the rows' prefixes, not positions."""
    hs, masker, wf = build(n, ne)
    a_idx, a_f = wf.sample_indices_gumbel(num, seed=123, compact_levels=False)
    b_idx, b_f = wf.sample_indices_gumbel(num, seed=123, compact_levels=True)
    pa, pb = torch.argsort(a_idx.view(-1)), torch.argsort(b_idx.view(-1))
    assert torch.equal(a_idx.view(-1)[pa], b_idx.view(-1)[pb])
    assert float((a_f[pa] - b_f[pb]).abs().max()) < 1e-14
    assert a_idx.shape[0] == min(num, a_idx.shape[0]) and abs(float(a_f.sum()) - 1.0) < 1e-12
    assert torch.unique(a_idx).shape[0] == a_idx.shape[0]


def test_count_splitting_at_full_size():
    """BASELINE config 4's size (36 qubits, 12 electrons, 1e7 samples): the counts add up to the number of samples drawn, every
    configuration is physical ((N, S_z) sector) and occurs once, the call is reproducible, and the sub-tree sharded sampler
    (emulated ranks) returns the same set."""
    from anqs_quantum_chemistry_b200 import dist as adist
    n, ne, num = 36, 12, 10 ** 7
    hs, masker, wf = build(n, ne)
    idx, cnt = wf.sample_stats(num, seed=5)
    assert float(cnt.real.sum()) == float(num) and float(cnt.imag.abs().max()) == 0.0 and float(cnt.real.min()) >= 1.0
    flat = idx.view(-1)
    even = torch.tensor(0x5555555555555555, dtype=torch.int64, device=DEV)
    assert bool((hs.popcount(flat & even) == ne // 2).all()) and bool((hs.popcount(flat & ~even) == ne // 2).all())
    assert torch.unique(flat).shape[0] == flat.shape[0] > 10 ** 6
    idx2, cnt2 = wf.sample_stats(num, seed=5)
    assert torch.equal(idx, idx2) and torch.equal(cnt, cnt2)
    parts = [adist.sharded_sample_stats(wf, num, seed=5, world_size=4, rank=r, gather=False) for r in range(4)]
    assert torch.equal(torch.cat([p[0] for p in parts]), idx) and torch.equal(torch.cat([p[1] for p in parts]), cnt)


@pytest.mark.parametrize('K', [1, 17, 1000, 100003])
def test_batch_reduce_gemm_matches_matmul(K):
    """k3_batch_reduce.cu (the batch reductions of the backward pass, torch's addmm backward under MLP:217-246) against
    float64 matmul on ragged problems: partial output tiles, odd leading dimensions, with and without column sums, accumulate."""
    from anqs_quantum_chemistry_b200 import _lib
    dev = torch.device('cuda:0')
    gen = torch.Generator(device='cpu').manual_seed(K)
    shapes = [(640, 64, 640, 64), (200, 56, 201, 57), (64, 64, 64, 64), (1, 1, 3, 5), (129, 7, 131, 7)]   # M, N, lda, ldb
    ops, problems = [], []
    for M, N, lda, ldb in shapes:
        A = torch.randn(K, lda, generator=gen, dtype=torch.float64).to(dev)
        Bm = torch.randn(K, ldb, generator=gen, dtype=torch.float64).to(dev)
        C = torch.full((M, N + 2), 7.0, dtype=torch.float64, device=dev)
        cs = torch.full((M,), 7.0, dtype=torch.float64, device=dev) if M != 64 else None
        ops.append((A, Bm, C, cs, M, N))
        problems.append((A.data_ptr(), lda, M, Bm.data_ptr(), ldb, N, C.data_ptr(), N + 2, cs.data_ptr() if cs is not None else None))
    _lib.batch_reduce(problems, K, False, dev)
    for A, Bm, C, cs, M, N in ops:
        ref = A[:, :M].T @ Bm[:, :N]
        scale = max(1.0, float(ref.abs().max()))
        assert float((C[:, :N] - ref).abs().max()) < 1e-12 * scale * max(1, K) ** 0.5
        assert bool((C[:, N:] == 7.0).all())                       # outside the problem: untouched
        if cs is not None:
            assert float((cs - A[:, :M].sum(0)).abs().max()) < 1e-12 * max(1, K) ** 0.5 * scale
    first = [op[2].clone() for op in ops]
    _lib.batch_reduce(problems, K, True, dev)                      # accumulate: exactly twice the first result (deterministic)
    for (A, Bm, C, cs, M, N), f in zip(ops, first):
        assert torch.equal(C[:, :N], 2 * f[:, :N])


@pytest.mark.parametrize('count', [1.0, 2.0, 37.0])
def test_split_level_draws_follow_the_conditional_distribution(count):
    """Exactness of the random mode of split_level_kernel (ANQS:557-591 draws a multinomial): 2e5 parents with the same
    conditional row and distinct keys; the histogram of their children against count * p by chi-square (63 - 6 masked outcomes
    => 56 degrees of freedom).  count = 1 exercises the single-draw path, 2 the inversion binomial, 37 both binomial samplers."""
    from anqs_quantum_chemistry_b200 import _lib
    import ctypes
    dev = torch.device('cuda:0')
    B, D = 200000, 64
    gen = torch.Generator().manual_seed(5)
    logits = torch.randn(D, generator=gen, dtype=torch.float64) * 1.5
    logits[[3, 17, 31, 32, 60, 63]] = -float('inf')                # masked outcomes (cond = -inf), incl. the last one
    p = torch.softmax(logits, 0)
    cond = (0.5 * torch.log(p)).repeat(B, 1).contiguous().to(dev)   # the kernel takes log|psi| conditionals: p = softmax(2 cond)
    counts = torch.full((B,), count, dtype=torch.float64, device=dev)
    memo = torch.zeros(B, dtype=torch.int32, device=dev)
    cont = torch.tensor([-1], dtype=torch.int64, device=dev)        # every continuation allowed by the symmetry table
    keys = (torch.arange(B, dtype=torch.int64, device=dev) * 2654435761 + 12345)
    child = torch.empty((B, D), dtype=torch.float64, device=dev)
    n_child = torch.empty(B, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().anqs_sampler_split_level(_lib.dptr(cond), D, 6, _lib.dptr(counts), _lib.dptr(memo), _lib.dptr(cont), 1, B,
                                                   3, 1, 99, 0, _lib.dptr(keys), _lib.dptr(child), _lib.dptr(n_child), _lib.dptr(None),
                                                   _lib.stream_ptr(dev)))
    # the byte path for single-sample parents names the same child as the dense row (and leaves the dense rows of others alone)
    child2 = torch.full((B, D), -5.0, dtype=torch.float64, device=dev)
    n_child2 = torch.empty(B, dtype=torch.int64, device=dev)
    single = torch.empty(B, dtype=torch.int8, device=dev)
    _lib.check(_lib.lib().anqs_sampler_split_level(_lib.dptr(cond), D, 6, _lib.dptr(counts), _lib.dptr(memo), _lib.dptr(cont), 1, B,
                                                   3, 1, 99, 0, _lib.dptr(keys), _lib.dptr(child2), _lib.dptr(n_child2), _lib.dptr(single),
                                                   _lib.stream_ptr(dev)))
    assert torch.equal(n_child, n_child2)
    if count == 1.0:
        assert torch.equal(single.long(), child.argmax(1)) and bool((child2 == -5.0).all())
    else:
        assert bool((single == -2).all()) and torch.equal(child, child2)
    child = child.cpu()
    assert bool((child.sum(1) == count).all())                      # every parent's samples are conserved
    assert bool((child == child.round()).all()) and bool((child >= 0).all())
    hist = child.sum(0)
    assert float(hist[p == 0].sum()) == 0.0                         # impossible outcomes never drawn
    expect = B * count * p[p > 0]
    chi2 = float(((hist[p > 0] - expect) ** 2 / expect).sum())
    assert chi2 < 57 + 6 * (2 * 57) ** 0.5, chi2                    # mean dof, six sigma
    assert int(n_child.sum()) == int((child > 0).sum())


def test_predicted_level_sizes_give_the_same_samples():
    """sample_stats sizes its levels from the previous call (one host read per call): same configurations and counts as the
    exact path that reads every level's size back, for the same seed; an under-predicted level falls back to the exact path."""
    hs, masker, wf = build(20, 14)
    N = 10 ** 6
    i0, c0 = wf.sample_stats(N, seed=7)                       # first call: exact, records the level sizes
    assert wf._level_hints[N][-1] == i0.shape[0]
    i1, c1 = wf.sample_stats(N, seed=7)                       # predicted capacities
    assert torch.equal(i0, i1) and torch.equal(c0, c1)
    i2, c2 = wf.sample_stats(N, seed=8)
    i3, c3 = wf.sample_stats(N, seed=8, exact_levels=True)
    assert torch.equal(i2, i3) and torch.equal(c2, c3)
    assert float(c2.real.sum()) == float(N)
    wf._level_hints[N] = [1] * len(wf._level_hints[N])        # hopeless prediction: every level overflows -> exact fallback
    wf._level_hints[N][0] = -10 ** 9
    i4, c4 = wf.sample_stats(N, seed=8)
    assert torch.equal(i4, i3) and torch.equal(c4, c3)
