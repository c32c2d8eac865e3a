"""Forward + backward of the fp64 transformer wave function on the C3 shape, 10^4 samples (for ncu and for timing)."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker, synthetic
from anqs_quantum_chemistry_b200.transformer_anqs import TransformerANQS, TransformerANQSConfig
dev = torch.device('cuda:0')
n, ne = 20, 14
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(1)
wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=2, head_num=4))
idx = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2, ne // 2, B, seed=1).view('int64')).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for r in range(reps + 1):
    wf.zero_grad()
    ev[0].record()
    lp = wf.log_psi_of_indices(idx)
    ev[1].record()
    lp.real.sum().backward()
    ev[2].record()
    torch.cuda.synchronize()
    print('forward %.3f ms, backward %.3f ms' % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
print('grad norm', float(wf.cat_grad.norm()))
