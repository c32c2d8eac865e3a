"""Small driver for profiling the tcgen05 MADE kernel: python scripts/bench_tc.py [batch]"""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry,
                                         LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, synthetic)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device('cuda:0')
hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=14),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
x = torch.from_numpy(synthetic.random_physical_samples(56, 7, 7, b, seed=1).view(np.int64)).to(dev)
with torch.no_grad():
    for fn, name in ((wf.log_psi_tc, 'tcgen05 tf32'), (wf.log_psi_of_indices, 'fp64')):
        for _ in range(3):
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name}: batch {b}: {ms:.3f} ms -> {b / ms * 1e3:.3e} amplitudes/s')
