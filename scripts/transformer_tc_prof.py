"""A few launches of the tcgen05 transformer kernel on the C3 shape (for ncu)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker, synthetic
from anqs_quantum_chemistry_b200.transformer_anqs import TransformerANQS, TransformerANQSConfig
dev = torch.device('cuda:0')
n, ne = 20, 14
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(1)
wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=2, head_num=4))
idx = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2, ne // 2, 10 ** 6, seed=1).view('int64')).to(dev).repeat(20)[: 1 << 18].contiguous()
with torch.no_grad():
    for _ in range(3):
        out = wf.log_psi_kernel(idx, precision='tf32')
torch.cuda.synchronize()
print('sum', complex(out.sum()))
