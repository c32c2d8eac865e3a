#!/usr/bin/env python
"""VMC iterations/s at 1/2/4/8 GPUs on the BASELINE config-5 shape (56 qubits, 14 e-, synthetic integrals with 8 irreps, MADE):
one iteration = sub-tree sharded count-splitting sampler (EXP:626-679's `sample`, ANQS:494-525) -> amplitudes of the local rows
with the autograd graph -> sharded sample-aware local energy (all-gather of (index, amplitude), PO:396-487 on this rank's rows,
all-reduce of the energy statistics) -> loss EXP:609 -> backward through the local rows -> ONE all-reduce of the flat gradient ->
Adam step (replicated parameters).  Strong scaling: the number of samples per iteration is fixed, the rows are split.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_vmc_sharded.py

One JSON line on rank 0; times are CUDA-event times, max over ranks."""
import argparse, json, os, sys, tempfile
# Batch sizes drift from one VMC iteration to the next (the number of unique samples, each rank's share of them): without size
# classes the caching allocator meets a slightly larger request every few iterations, cannot reuse the cached block and
# calls cudaMalloc again - measured on 4 GPUs: one rank grew from 7 to 15 GiB reserved in ten iterations and single
# iterations took 35-195 ms instead of 18.  Must be set before the first CUDA allocation.
os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'roundup_power2_divisions:8')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry,
                                         SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, synthetic)
from anqs_quantum_chemistry_b200 import dist as adist

ap = argparse.ArgumentParser()
ap.add_argument('--samples', type=int, default=10 ** 6)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--qubits', type=int, default=56)
ap.add_argument('--electrons', type=int, default=14)
ap.add_argument('--profile', action='store_true', help='after the timed run: one iteration under the torch profiler (rank 0 prints the tables)')
ap.add_argument('--f64-sampler', action='store_true', help='conditional probabilities of the sampler in float64 instead of tf32')
args = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n, n_el = args.qubits, args.electrons
xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=8, seed=0)
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix=f'anqs_vmc_r{rank}_'), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
if not args.f64_sampler:
    wf.set_inference_precision('tf32')          # only the sampler's conditionals: amplitudes with a graph stay float64
ham.tables
opt = torch.optim.Adam(wf.parameters(), lr=1e-3)
sle = adist.ShardedLocalEnergy(ham, n_el // 2, n_el // 2)
grad_step = adist.ShardedEnergyGradient(wf, sle.stats)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def iteration(it, marks=None):
    def mark():
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)
    mark()
    idx, cnt = adist.sharded_sample_stats(wf, args.samples, seed=1000 + it, gather=False)
    mark()
    mean, var, loss = grad_step(idx)
    mark()
    opt.step()
    mark()
    return idx.shape[0], mean


for it in range(args.warmup):
    iteration(it)
    if it == 0:
        adist.reserve_device_memory(dev)   # no cudaMalloc in the iterations that follow (dist.py)
barrier()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
phase = torch.zeros(3, dtype=torch.float64)
start.record()
all_marks, rows = [], 0
for it in range(args.steps):
    marks = []
    n_rows, mean = iteration(args.warmup + it, marks)
    all_marks.append(marks)
    rows += n_rows
    if os.environ.get('ANQS_ALLOC_TRACE'):
        torch.cuda.synchronize()
        st = torch.cuda.memory_stats()
        print(f"rank {rank} iteration {it}: rows {n_rows} sampler {marks[0].elapsed_time(marks[1]):.1f} ms step {marks[1].elapsed_time(marks[2]):.1f} ms "
              f"cudaMalloc {st.get('num_device_alloc')} cudaFree {st.get('num_device_free')} retries {st.get('num_alloc_retries')} "
              f"reserved {st.get('reserved_bytes.all.current', 0) / 2**30:.2f} GiB", file=sys.stderr, flush=True)
stop.record()
barrier()
for marks in all_marks:
    for k in range(3):
        phase[k] += marks[k].elapsed_time(marks[k + 1])
if rank == 0:
    print('per-iteration ms (rank 0):', ' '.join(f'{m[0].elapsed_time(m[3]):.1f}' for m in all_marks), file=sys.stderr, flush=True)
t = torch.tensor([start.elapsed_time(stop)] + (phase / args.steps).tolist() + [0.0], dtype=torch.float64, device=dev)
r = torch.tensor([float(rows)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(r)
if rank == 0:
    ms = float(t[0]) / args.steps
    print(json.dumps({'metric': 'vmc_iterations_per_sec', 'value': 1e3 / ms, 'unit': 'it/s', 'n_gpus': world, 'steps': args.steps,
                      'warmup': args.warmup, 'ms_per_iteration': ms, 'scaling': 'strong',
                      'config': {'workload': f'C5 shape: {n} qubits, {n_el} e-, synthetic integrals with 8 irreps, MADE ({wf.param_num} parameters)',
                                 'samples_per_iteration': args.samples, 'unique_per_iteration': float(r) / args.steps,
                                 'terms': int(ham.term_num), 'unique_xy_masks': int(ham.unq_xy_masks_num),
                                 'sampler_conditionals': 'f64' if args.f64_sampler else 'tf32 (tcgen05)',
                                 'amplitudes_and_gradient': 'f64', 'optimizer': 'Adam'},
                      'phase_ms_max_over_ranks': {'sampler': float(t[1]), 'amplitudes+local_energy+backward+all_reduce': float(t[2]),
                                                  'optimizer': float(t[3])},
                      'unique_rows_per_s': float(r) / args.steps / ms * 1e3,
                      'energy_mean_last': [float(mean.real), float(mean.imag)]}), flush=True)
if args.profile:
    from torch.profiler import profile, ProfilerActivity
    barrier()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        iteration(10 ** 6)
        torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=60))
        print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=60))
if world > 1:
    dist.destroy_process_group()
