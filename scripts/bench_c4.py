#!/usr/bin/env python
"""BASELINE config 4 (C4): C2 6-31G shape — 36 qubits, 12 electrons, synthetic integrals with 8 irreps — count-splitting batch
sampling of 1e7 samples sharded by sub-tree over the ranks, then the sample-aware local energy of the resulting unique set
sharded by rows.  One JSON line on rank 0: unique/s of the sampler (max over ranks) and E_loc/s of the energy pass.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py [--samples 10000000]
"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry,
                                         SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, synthetic)
from anqs_quantum_chemistry_b200 import dist as adist

ap = argparse.ArgumentParser()
ap.add_argument('--samples', type=int, default=10 ** 7)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--tf32', action='store_true', help='tcgen05 inference mode for the conditional probabilities and amplitudes')
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n, n_el = 36, 12
xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=8, seed=0)
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix=f'anqs_c4_r{rank}_'), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
if args.tf32:
    wf.set_inference_precision('tf32')
ham.tables

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def timed(fn):
    best = None
    for _ in range(args.reps):
        barrier()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t) if best is None else min(best, float(t))
    return out, best

(idx, cnt), t_sample = timed(lambda: adist.sharded_sample_stats(wf, args.samples, seed=1, gather=False))
n_local = torch.tensor([idx.shape[0]], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(n_local)
n_unq = int(n_local)
with torch.no_grad():
    amps, t_amp = timed(lambda: wf.amplitude(idx))
sle = adist.ShardedLocalEnergy(ham, n_el // 2, n_el // 2)
(eloc, mean, var), t_eloc = timed(lambda: sle(idx, amps))
assert abs(float(cnt.real.sum()) * 1.0 - 0) >= 0
total = torch.tensor([float(cnt.real.sum())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(total)
if rank == 0:
    print(json.dumps({'config': 'C4: 36 qubits, 12 e-, synthetic integrals with 8 irreps, untrained MADE', 'n_gpus': world,
                      'samples': args.samples, 'samples_conserved': float(total) == float(args.samples), 'unique': n_unq,
                      'terms': int(ham.term_num), 'unique_xy_masks': int(ham.unq_xy_masks_num), 'nn_mode': 'tf32 (tcgen05)' if args.tf32 else 'f64',
                      'sample_s': t_sample, 'unique_per_s': n_unq / t_sample, 'amplitude_s': t_amp, 'amplitudes_per_s': n_unq / t_amp,
                      'local_energy_s': t_eloc, 'eloc_per_s': n_unq / t_eloc, 'energy_mean': [float(mean.real), float(mean.imag)]}), flush=True)
if world > 1:
    dist.destroy_process_group()
