"""Kernel-level breakdown of the amplitude forward (with graph) + backward at the C5 shape, ~1e6 rows (torch profiler table)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         LogAbsPhaseANQS, ANQSConfig, synthetic)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mode = sys.argv[2] if len(sys.argv) > 2 else 'MADE'
dev = torch.device('cuda:0')
hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=14),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode=mode))
idx = torch.from_numpy(synthetic.random_physical_samples(56, 7, 7, rows, seed=1).view(np.int64)).to(dev).view(-1, 1)
seed = torch.randn(rows, dtype=torch.complex128, device=dev)
def step():
    for p in wf.parameters():
        p.grad = None
    amps = wf.amplitude(idx)
    loss = 2 * (seed * torch.log(torch.conj(amps))).sum().real
    loss.backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for p in wf.parameters():
    p.grad = None
ev[0].record(); amps = wf.amplitude(idx); loss = 2 * (seed * torch.log(torch.conj(amps))).sum().real; ev[1].record(); loss.backward(); ev[2].record()
torch.cuda.synchronize()
print(f'rows {rows} {mode}: forward {ev[0].elapsed_time(ev[1]):.2f} ms, backward {ev[1].elapsed_time(ev[2]):.2f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=70))
