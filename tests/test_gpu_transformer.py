"""GPU parity: the transformer wave-function kernel (k5_transformer.cu, through TransformerANQS and the C ABI) against
(a) the logits of the reference's own TransformerMADE module + the numpy masking oracle and (b) the torch module it mirrors.
fp64: 1e-10 on log psi and the conditionals; samplers checked by their invariants."""
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         TransformerANQS, TransformerANQSConfig, synthetic)
from oracle import anqs_numpy as onp

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
CASES = ['tfm_n12', 'tfm_n14', 'tfm_n20']


def build(n, ne, depth, head_num, seed):
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=tempfile.mkdtemp(prefix='anqs_tfm_test_'), rng_seed=0)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=depth, head_num=head_num))
    # the network itself is re-created on the CPU under the golden's seed (device-side init would draw other numbers)
    from anqs_quantum_chemistry_b200.transformer_anqs import TransformerMADE
    torch.manual_seed(seed)
    cpu_net = TransformerMADE(dim=64, out_dim=4, depth=depth, qubit_num=n, head_num=head_num, dtype=torch.float64)
    wf.transformer_made.load_state_dict(cpu_net.state_dict())
    return hs, masker, wf


def case(name):
    g = load_golden(name)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    hs, masker, wf = build(n, ne, int(g['depth']), int(g['head_num']), int(g['seed']))
    return g, onp.NumberSpinMasks(n, ne, qubit_per_qudit=1), wf


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('name', CASES)
def test_log_psi_matches_reference_logits(name):
    g, masks, wf = case(name)
    x = g['samples'].view(np.uint64)
    ref = onp.transformer_log_psi_from_logits(g['logits'], x, masks)
    lp = wf.log_psi_kernel(_dev(g['samples'])).cpu().numpy()
    nphys = int(g['n_phys'])
    assert np.array_equal(np.isneginf(lp.real), np.isneginf(ref.real))
    assert np.abs(lp[:nphys] - ref[:nphys]).max() < 1e-10
    with torch.no_grad():
        amp = wf.amplitude(_dev(g['samples']).view(-1, 1)).cpu().numpy()
    assert np.abs(amp - np.exp(ref)).max() < 1e-10


@pytest.mark.parametrize('name', CASES)
def test_conditionals_match_reference_logits(name):
    g, masks, wf = case(name)
    n, nphys = masks.n, int(g['n_phys'])
    x = g['samples'].view(np.uint64)[:nphys]
    for t in (0, 1, n // 2, n - 1):
        logits = (g['prefix_logits'] if t == n // 2 else g['logits'])[:nphys]
        ref = onp.transformer_cond_from_logits(logits, x, t, masks)
        c = wf.cond_log_abs(qudit_idx=t, prefix_idx=_dev(x.view(np.int64))).cpu().numpy()
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < 1e-10


@pytest.mark.parametrize('name', CASES)
def test_kernel_matches_torch_path_and_backward_kernel_matches_autograd(name):
    """Forward: k5_transformer.cu against the torch module.  Backward: k5_transformer_bwd.cu + k3_batch_reduce.cu (what
    log_psi_of_indices differentiates through) against autograd through the torch module, every parameter, 1e-10 relative
    to the largest gradient entry; per-sample weights on both log|psi| and arg psi so that no term cancels."""
    import anqs_quantum_chemistry_b200.transformer_anqs as tfm
    g, masks, wf = case(name)
    n, ne = masks.n, masks.particle_num
    B = 203
    x = _dev(synthetic.random_physical_samples(n, ne // 2, ne // 2, B, seed=5).view(np.int64)).view(-1, 1)
    gen = torch.Generator().manual_seed(11)
    a, b = torch.randn(B, generator=gen, dtype=torch.float64).to(DEV), torch.randn(B, generator=gen, dtype=torch.float64).to(DEV)
    lp_k = wf.log_psi_kernel(x)
    lp_t = wf.log_psi_torch(x)
    assert (lp_k - lp_t.detach()).abs().max() < 1e-10
    wf.zero_grad()
    (a * lp_t.real + b * lp_t.imag).sum().backward()
    ref = {k: p.grad.clone() for k, p in wf.named_parameters()}
    for scratch in (8 * 2 ** 20, tfm._BWD_SCRATCH_BYTES):               # several chunks that accumulate; one chunk
        old, tfm._BWD_SCRATCH_BYTES = tfm._BWD_SCRATCH_BYTES, scratch
        try:
            wf.zero_grad()
            lp_c = wf.log_psi_of_indices(x)        # grad mode on -> _TransformerLogPsi
            assert lp_c.requires_grad and torch.equal(lp_c.detach(), lp_k)
            (a * lp_c.real + b * lp_c.imag).sum().backward()
        finally:
            tfm._BWD_SCRATCH_BYTES = old
        for k, p in wf.named_parameters():
            scale = float(ref[k].abs().max()) + 1e-300
            assert float((p.grad - ref[k]).abs().max()) <= 1e-10 * max(scale, 1.0), (k, float((p.grad - ref[k]).abs().max()), scale)
    g2 = wf.cat_grad
    assert g2.shape[0] == wf.param_num and bool(torch.isfinite(g2).all()) and float(g2.abs().max()) > 0
    # the same call (one chunk) twice gives the same bits (fixed-order reductions)
    first = g2.clone()
    wf.zero_grad()
    lp_c = wf.log_psi_of_indices(x)
    (a * lp_c.real + b * lp_c.imag).sum().backward()
    assert torch.equal(wf.cat_grad, first)
    # two forward calls before the first backward: the first one's activations are gone, its backward recomputes them
    wf.zero_grad()
    lp_1 = wf.log_psi_of_indices(x)
    lp_2 = wf.log_psi_of_indices(x[:50])
    (a * lp_1.real + b * lp_1.imag).sum().backward()
    for k, p in wf.named_parameters():
        scale = max(float(ref[k].abs().max()), 1.0)
        assert float((p.grad - ref[k]).abs().max()) <= 1e-10 * scale, k
    wf.zero_grad()
    (a[:50] * lp_2.real + b[:50] * lp_2.imag).sum().backward()     # this one still owns the workspace
    lp_t = wf.log_psi_torch(x[:50])
    g_own = wf.cat_grad.clone()
    wf.zero_grad()
    (a[:50] * lp_t.real + b[:50] * lp_t.imag).sum().backward()
    assert float((g_own - wf.cat_grad).abs().max()) <= 1e-10 * max(float(g_own.abs().max()), 1.0)
    # tile boundaries: any batch size gives the same numbers
    for bsz in (1, 2, 3, 4, 61):
        assert torch.equal(wf.log_psi_kernel(x[:bsz]), lp_k[:bsz])
        assert torch.equal(wf.log_psi_of_indices(x[:bsz]).detach(), lp_k[:bsz])
    assert wf.log_psi_kernel(x[:0]).shape[0] == 0
    assert wf.log_psi_of_indices(x[:0]).shape[0] == 0


@pytest.mark.parametrize('name', CASES)
def test_backward_kernel_matches_reference_module_gradient(name):
    """Gradient golden (oracle/make_golden.py: autograd through the REFERENCE's TransformerMADE + the masked normalisation) of
    sum_i a_i log|psi(x_i)| + b_i arg psi(x_i): every parameter tensor of k5_transformer_bwd.cu's result within 1e-10."""
    g, masks, wf = case(name)
    gg = load_golden(name + '_grad')
    nphys = int(g['n_phys'])
    x = _dev(g['samples'][:nphys]).view(-1, 1)
    a, b = _dev(gg['a']), _dev(gg['b'])
    wf.zero_grad()
    lp = wf.log_psi_of_indices(x)
    assert float((lp.real - _dev(gg['log_abs'])).abs().max()) < 1e-10 and float((lp.imag - _dev(gg['phase'])).abs().max()) < 1e-10
    (a * lp.real + b * lp.imag).sum().backward()
    names = [str(k) for k in g['param_names']]
    ours = dict(wf.transformer_made.named_parameters())
    assert list(ours) == names
    for i, k in enumerate(names):
        ref = _dev(gg[f'grad_{i:02d}'])
        err = float((ours[k].grad - ref).abs().max())
        assert err <= 1e-10 * max(1.0, float(ref.abs().max())), (k, err)


def test_normalisation_and_samplers():
    g, masks, wf = case('tfm_n12')
    allx = torch.arange(2 ** 12, dtype=torch.int64, device=DEV).view(-1, 1)
    with torch.no_grad():
        p = (wf.amplitude(allx).abs() ** 2).cpu().numpy()
    assert abs(p.sum() - 1.0) < 1e-12 and (p > 0).sum() == 225
    idx, cnt = wf.sample_stats(10 ** 6, seed=3)
    c = cnt.real
    assert float(c.sum()) == 1e6 and 100 < idx.shape[0] <= 225 and float(c.min()) >= 1.0
    assert bool((p[idx.view(-1).cpu().numpy()] > 0).all())                     # every sampled configuration is physical
    with torch.no_grad():
        expect = (wf.amplitude(idx).abs() ** 2) * 1e6
    big = expect > 20
    chi2 = float((((c - expect) ** 2) / expect)[big].sum() / big.sum())
    assert 0.6 < chi2 < 1.5, chi2
    gi, gf = wf.sample_indices_gumbel(100, seed=4)
    assert gi.shape[0] == 100 and gi.view(-1).unique().shape[0] == 100 and abs(float(gf.sum()) - 1.0) < 1e-12


def test_vmc_iterations_lower_the_energy():
    """Config 3 end to end at a small size: transformer ansatz, Gumbel unique sampling, sample-aware local energies, loss
    EXP:609, Adam.  The energy after 30 iterations is below the first one and above the sector's ground state."""
    from anqs_quantum_chemistry_b200 import (PauliObservable, PauliArraysOperator, SamplingConfig, SamplingResult, sample,
                                             LocalEnergyCalculationConfig, compute_local_energies, vmc_loss)
    n, ne = 12, 4
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
    hs, masker, wf = build(n, ne, depth=2, head_num=4, seed=1)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    opt = torch.optim.Adam(wf.parameters(), lr=3e-3)
    energies = []
    for it in range(30):
        opt.zero_grad()
        res, _, _, _ = sample(wf=wf, config=SamplingConfig(sample_indices=True, sample_num=225), seed=100 + it)
        indices, perm = wf.sort_base_idx(res.indices)
        amps = wf.amplitude(indices)
        le, _ = compute_local_energies(wf=wf, sampling_result=SamplingResult(indices=indices, counts=res.counts[perm]),
                                       sampled_amps=amps.detach(), ham=ham,
                                       config=LocalEnergyCalculationConfig(use_tree_for_candidates='ham'), sample_aware=True)
        est = le.sample_aware_e_loc_mc_est
        vmc_loss(amps, est).backward()
        opt.step()
        energies.append(float(est.mean.real))
    assert energies[-1] < energies[0] - 1e-3
    assert abs(float(est.mean.imag)) < 1e-9


@pytest.mark.parametrize('n,ne,heads,depth', [(12, 4, 4, 2), (20, 14, 4, 2), (20, 14, 8, 1), (14, 10, 16, 3)])
def test_tensor_core_mode_within_tolerance(n, ne, heads, depth):
    """tcgen05 (tf32) inference mode against the fp64 kernel: the same symmetry masks exactly (an unphysical configuration is
    -inf in both), log|psi| within 2e-2 and the phase within 5e-2 rad (tf32 products carry 10 mantissa bits; measured 3e-3 /
    7e-3), conditionals within 1e-2, and sampling through it stays physical and normalised."""
    from anqs_quantum_chemistry_b200 import synthetic
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=tempfile.mkdtemp(prefix='anqs_tfm_tc_'), rng_seed=0)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(1)
    wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=depth, head_num=heads))
    phys = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2, ne // 2, 3000, seed=1).view('int64')).to(DEV)
    unphys = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2 + 1, ne // 2 - 1, 50, seed=2).view('int64')).to(DEV)
    idx = torch.cat((phys, unphys))
    with torch.no_grad():
        ref = wf.log_psi_kernel(idx, precision='fp64')
        tc = wf.log_psi_kernel(idx, precision='tf32')
    fin = torch.isfinite(ref.real)
    assert torch.equal(fin, torch.isfinite(tc.real)) and int(fin.sum()) == phys.shape[0]
    assert float((tc.real[fin] - ref.real[fin]).abs().max()) < 2e-2
    assert float((tc.imag[fin] - ref.imag[fin]).abs().max()) < 5e-2
    for q in (0, n // 2, n - 1):
        wf.set_inference_precision('fp64')
        c0 = wf.cond_log_abs(qudit_idx=q, prefix_idx=phys)
        wf.set_inference_precision('tf32')
        c1 = wf.cond_log_abs(qudit_idx=q, prefix_idx=phys)
        f = torch.isfinite(c0)
        assert torch.equal(f, torch.isfinite(c1))
        assert float((c0[f] - c1[f]).abs().max()) < 1e-2
    # samplers on the tensor-core conditionals
    s_idx, s_cnt = wf.sample_stats(10 ** 5, seed=3)
    assert float(s_cnt.real.sum()) == 1e5
    even = torch.tensor(0x5555555555555555, dtype=torch.int64, device=DEV)
    assert bool((hs.popcount(s_idx.view(-1) & even) == ne // 2).all()) and bool((hs.popcount(s_idx.view(-1) & ~even) == ne // 2).all())
    g_idx, g_f = wf.sample_indices_gumbel(min(200, phys.shape[0]))
    assert abs(float(g_f.sum()) - 1.0) < 1e-12 and torch.unique(g_idx).shape[0] == g_idx.shape[0]
    # a parameter update is picked up (the packed weights are rebuilt)
    with torch.no_grad():
        for p in wf.parameters():
            p.add_(0.01 * torch.randn_like(p))
        ref2 = wf.log_psi_kernel(phys, precision='fp64')
        tc2 = wf.log_psi_kernel(phys, precision='tf32')
    assert float((tc2.real - ref2.real).abs().max()) < 2e-2 and float((ref2.real - ref.real[fin]).abs().max()) > 1e-3
