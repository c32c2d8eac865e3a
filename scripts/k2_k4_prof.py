"""Kernel families 2 and 4 on the C5 shape for ncu: table build over 2^20 sampled configurations, count-splitting sampler of 1e6 samples."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         LogAbsPhaseANQS, ANQSConfig, synthetic)
from anqs_quantum_chemistry_b200.hilbert_space import SampleTable
dev = torch.device('cuda:0')
hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=14),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
wf.set_inference_precision('tf32')
n = 1 << 20
idx = torch.from_numpy(synthetic.random_physical_samples(56, 7, 7, n, seed=1).view(np.int64)).to(dev)
amps = torch.from_numpy(synthetic.random_amplitudes(n, seed=2)).to(dev)
for _ in range(2):
    t = SampleTable(idx, amps)
    wf.sample_stats(10 ** 6, seed=3)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); t = SampleTable(idx, amps); e1.record(); i, c = wf.sample_stats(10 ** 6, seed=4); e2.record(); torch.cuda.synchronize()
print(f'table build over {n} keys {e0.elapsed_time(e1):.3f} ms; sample_stats(1e6) -> {i.shape[0]} unique {e1.elapsed_time(e2):.3f} ms')
