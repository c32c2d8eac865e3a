"""NADE mode (the reference's default): fp64 kernel vs tcgen05 tf32 kernel at 56 qubits - amplitudes/s, sample_stats(1e6), errors."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker,
                                         LogAbsPhaseANQS, ANQSConfig, synthetic)
dev = torch.device('cuda:0')
n, ne = 56, 14
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='NADE'))
idx = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2, ne // 2, 1 << 18, seed=1).view('int64')).to(dev)
res = {}
with torch.no_grad():
    for prec in ('fp64', 'tf32'):
        wf.set_inference_precision(prec)
        res[prec] = wf.log_psi_of_indices(idx); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): wf.log_psi_of_indices(idx)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        wf.sample_stats(10 ** 6, seed=1); torch.cuda.synchronize(); t0 = time.perf_counter()
        si, sc = wf.sample_stats(10 ** 6, seed=1); torch.cuda.synchronize(); ds = time.perf_counter() - t0
        print(f'{prec}: {idx.shape[0] / dt:.3e} amplitudes/s; sample_stats(1e6) -> {si.shape[0]} unique in {ds * 1e3:.1f} ms')
print(f'max |d log|psi|| = {float((res["fp64"].real - res["tf32"].real).abs().max()):.2e}, max |d phase| = {float((res["fp64"].imag - res["tf32"].imag).abs().max()):.2e}')
