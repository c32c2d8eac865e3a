"""Kernel-level breakdown of the count-splitting sampler at the C5 shape (1e6 samples, MADE, tf32 conditionals)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         LogAbsPhaseANQS, ANQSConfig)
from anqs_quantum_chemistry_b200 import dist as adist
samples = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 6
dev = torch.device('cuda:0')
hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=14),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
wf.set_inference_precision(sys.argv[2] if len(sys.argv) > 2 else 'tf32')
for it in range(3):
    idx, cnt = adist.sharded_sample_stats(wf, samples, seed=it, gather=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); idx, cnt = adist.sharded_sample_stats(wf, samples, seed=7, gather=False); e1.record(); torch.cuda.synchronize()
print(f'{samples} samples -> {idx.shape[0]} unique in {e0.elapsed_time(e1):.2f} ms')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    adist.sharded_sample_stats(wf, samples, seed=8, gather=False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=80))
