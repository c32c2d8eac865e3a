#!/usr/bin/env python
"""2+ GPUs under torchrun: the sharded VMC gradient (sub-tree sharded sampler -> local amplitudes -> sharded local energy ->
local backward -> one all-reduce) equals the gradient rank 0 computes alone on the gathered batch.  Prints one line on rank 0."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry, SpinHalfProjectionSymmetry,
                                         LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, synthetic, MonteCarloEstimator, vmc_loss)
from anqs_quantum_chemistry_b200 import dist as adist
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n, n_el = 20, 14
xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix=f'anqs_sg_r{rank}_'), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
idx, cnt = adist.sharded_sample_stats(wf, 10 ** 6, seed=1, gather=False)
sle = adist.ShardedLocalEnergy(ham, n_el // 2, n_el // 2)
step = adist.ShardedEnergyGradient(wf, sle.stats)
for _ in range(2):
    mean, var, loss = step(idx)
torch.cuda.synchronize()
t0 = time.perf_counter()
mean, var, loss = step(idx)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
g_sharded = torch.cat([p.grad.reshape(-1) for p in wf.parameters()]).clone()
# the same on one GPU over the gathered batch
g_idx, _ = adist.sharded_sample_stats(wf, 10 ** 6, seed=1, gather=True)
for p in wf.parameters():
    p.grad = None
amps = wf.amplitude(g_idx)
e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=g_idx, unq_batch_as_amps=amps.detach(), coupling_method='ham',
                                             alpha_num=n_el // 2, beta_num=n_el // 2)
a = amps.detach()
est = MonteCarloEstimator(values=e, counts=a.conj() * a)
ref_loss = vmc_loss(amps, est)
ref_loss.backward()
g_ref = torch.cat([p.grad.reshape(-1) for p in wf.parameters()])
err = float((g_sharded - g_ref).abs().max())
scale = float(g_ref.abs().max())
if rank == 0:
    print(f'world={world} unique={g_idx.shape[0]} (local {idx.shape[0]}) energy={complex(mean):.12f} vs {complex(est.mean):.12f} '
          f'loss {float(loss):.3e} vs {float(ref_loss):.3e}; max|g_sharded - g_single| = {err:.3e} (scale {scale:.3e}); sharded step {dt * 1e3:.2f} ms')
assert err < 1e-10 * max(1.0, scale) and abs(complex(mean) - complex(est.mean)) < 1e-10
if world > 1:
    dist.destroy_process_group()
