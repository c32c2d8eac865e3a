"""Phase timing of one VMC iteration on the C3 shape (20 qubits, N_unq = 1e4) — where the 4 ms go."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry,
                                         SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig,
                                         SamplingConfig, SamplingResult, sample, LocalEnergyCalculationConfig,
                                         compute_local_energies, vmc_loss, synthetic)
dev = torch.device('cuda:0')
n, n_el, sample_num = 20, 14, 10 ** 4
xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix='anqs_vmc_'), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
opt = torch.optim.Adam(wf.parameters(), lr=1e-3)
cfg_s = SamplingConfig(sample_indices=True, sample_num=sample_num)
cfg_e = LocalEnergyCalculationConfig(use_tree_for_candidates='ham')
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    return t
def one_iter(timed):
    t = time.perf_counter()
    opt.zero_grad()
    if timed: t = tick('zero_grad', t)
    res, _, _, _ = sample(wf=wf, config=cfg_s)
    if timed: t = tick('sample (gumbel)', t)
    indices, perm = wf.sort_base_idx(res.indices)
    if timed: t = tick('sort', t)
    amps = wf.amplitude(indices)
    if timed: t = tick('amplitude fwd', t)
    le, _ = compute_local_energies(wf=wf, sampling_result=SamplingResult(indices=indices, counts=res.counts[perm]),
                                   sampled_amps=amps.detach(), ham=ham, config=cfg_e, sample_aware=True)
    if timed: t = tick('local energies', t)
    est = le.sample_aware_e_loc_mc_est
    loss = vmc_loss(amps, est)
    if timed: t = tick('loss', t)
    loss.backward()
    if timed: t = tick('backward', t)
    opt.step()
    if timed: t = tick('adam', t)
for _ in range(3): one_iter(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): one_iter(False)
torch.cuda.synchronize()
print('untimed: %.3f ms/iter' % ((time.perf_counter() - t0) * 50))
for _ in range(20): one_iter(True)
for k, v in acc.items(): print(f'  {k:18s} {v * 50:.3f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): one_iter(False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=30, max_name_column_width=60))
