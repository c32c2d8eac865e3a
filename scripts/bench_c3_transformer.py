"""BASELINE config 3 (C3): 20 qubits, 14 electrons, dense synthetic H, transformer wave function (dim 64, depth 2, 4 heads) with
particle-number / S_z masks: amplitudes/s of the hand-written kernel, unique samples/s of both samplers, VMC iterations/s."""
import sys, os, tempfile, time, json
# Batch sizes drift from one VMC iteration to the next (the number of unique samples, each rank's share of them): without size
# classes the caching allocator meets a slightly larger request every few iterations, cannot reuse the cached block and
# calls cudaMalloc again - measured on 4 GPUs: one rank grew from 7 to 15 GiB reserved in ten iterations and single
# iterations took 35-195 ms instead of 18.  Must be set before the first CUDA allocation.
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'roundup_power2_divisions:8')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry, SpinHalfProjectionSymmetry,
                                         LocallyDecomposableMasker, SamplingConfig, SamplingResult, sample, LocalEnergyCalculationConfig,
                                         compute_local_energies, vmc_loss, synthetic)
from anqs_quantum_chemistry_b200.transformer_anqs import TransformerANQS, TransformerANQSConfig
dev = torch.device('cuda:0')
n, n_el = 20, 14
xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix='anqs_c3_'), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=2, head_num=4))
out = {'config': 'C3: 20 qubits, 14 e-, dense synthetic H (T = %d), TransformerANQS dim 64 depth 2 heads 4' % ham.term_num}

def wall(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) / reps

sector = torch.from_numpy(synthetic.random_physical_samples(n, n_el // 2, n_el // 2, 10 ** 6, seed=1).view('int64')).to(dev)
with torch.no_grad():
    big = sector.repeat(20)[: 1 << 18].contiguous()
    for prec in ('fp64', 'tf32'):
        wf.set_inference_precision(prec)
        _, t = wall(lambda: wf.log_psi_kernel(big))
        out[f'amplitudes_per_s_kernel_{prec}'] = big.shape[0] / t
        (idx, cnt), t = wall(lambda: wf.sample_stats(10 ** 7, seed=1), reps=3, warm=1)
        out[f'count_splitting_1e7_samples_{prec}'] = {'unique': int(idx.shape[0]), 'seconds': t, 'unique_per_s': idx.shape[0] / t}
        (idx, f), t = wall(lambda: wf.sample_indices_gumbel(10 ** 4), reps=5, warm=2)
        out[f'gumbel_1e4_{prec}'] = {'unique': int(idx.shape[0]), 'seconds': t}
    ref = wf.log_psi_kernel(sector[:5000], precision='fp64'); tc = wf.log_psi_kernel(sector[:5000], precision='tf32')
    out['tf32_max_abs_err'] = {'log_abs': float((ref.real - tc.real).abs().max()), 'phase': float((ref.imag - tc.imag).abs().max())}
wf.set_inference_precision('tf32')   # the VMC iteration below samples through the tensor-core conditionals
opt = torch.optim.Adam(wf.parameters(), lr=1e-3)
cfg_s, cfg_e = SamplingConfig(sample_indices=True, sample_num=10 ** 4), LocalEnergyCalculationConfig(use_tree_for_candidates='ham')
energies = []
def one_iter():
    opt.zero_grad()
    res, _, _, _ = sample(wf=wf, config=cfg_s)
    indices, perm = wf.sort_base_idx(res.indices)
    amps = wf.amplitude(indices)
    le, _ = compute_local_energies(wf=wf, sampling_result=SamplingResult(indices=indices, counts=res.counts[perm]), sampled_amps=amps.detach(),
                                   ham=ham, config=cfg_e, sample_aware=True)
    est = le.sample_aware_e_loc_mc_est
    vmc_loss(amps, est).backward()
    opt.step()
    energies.append(float(est.mean.real))
_, t = wall(one_iter, reps=20, warm=3)
out['vmc_iteration'] = {'ms_per_iter': t * 1e3, 'iters_per_s': 1 / t, 'n_unq': 10 ** 4, 'energy_first': energies[0], 'energy_last': energies[-1],
                        'note': 'sampler on the tcgen05 conditionals, E_loc on the fused kernel, gradient through k5_transformer_bwd.cu + k3_batch_reduce.cu'}
# where the time of an iteration goes
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
res, _, _, _ = sample(wf=wf, config=cfg_s)
indices, perm = wf.sort_base_idx(res.indices)
parts = {'sample_ms': timed(lambda: sample(wf=wf, config=cfg_s)), 'sort_ms': timed(lambda: wf.sort_base_idx(res.indices))}
def fwd_bwd():
    opt.zero_grad()
    a = wf.amplitude(indices)
    (a.log_psi.real.sum()).backward()
parts['amplitude_forward_backward_ms'] = timed(fwd_bwd)
with torch.no_grad():
    amps = wf.amplitude(indices)
parts['local_energy_ms'] = timed(lambda: compute_local_energies(wf=wf, sampling_result=SamplingResult(indices=indices, counts=res.counts[perm]),
                                                                sampled_amps=amps, ham=ham, config=cfg_e, sample_aware=True))
out['vmc_iteration']['parts'] = parts
print(json.dumps(out))
