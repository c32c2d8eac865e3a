"""tcgen05 transformer kernels against the fp64 kernels: max errors of log|psi|, phase and the conditionals; throughput."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker, synthetic
from anqs_quantum_chemistry_b200.transformer_anqs import TransformerANQS, TransformerANQSConfig
dev = torch.device('cuda:0')
for n, ne, heads, depth in ((12, 4, 4, 2), (20, 14, 4, 2), (20, 14, 8, 1), (14, 10, 16, 3)):
    hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(1)
    wf = TransformerANQS(hilbert_space=hs, masker=masker, config=TransformerANQSConfig(dim=64, depth=depth, head_num=heads))
    idx = torch.from_numpy(synthetic.random_physical_samples(n, ne // 2, ne // 2, 5000, seed=1).view('int64')).to(dev)
    with torch.no_grad():
        ref = wf.log_psi_kernel(idx, precision='fp64')
        tc = wf.log_psi_kernel(idx, precision='tf32')
        torch.cuda.synchronize()
        print(f'n={n} heads={heads} depth={depth}: {idx.shape[0]} samples, max |d log|psi|| = {float((tc.real - ref.real).abs().max()):.3e}, '
              f'max |d phase| = {float((tc.imag - ref.imag).abs().max()):.3e}  (|log psi| up to {float(ref.real.abs().max()):.2f})', flush=True)
        for q in (0, n // 2, n - 1):
            wf.set_inference_precision('fp64'); c0 = wf.cond_log_abs(qudit_idx=q, prefix_idx=idx)
            wf.set_inference_precision('tf32'); c1 = wf.cond_log_abs(qudit_idx=q, prefix_idx=idx)
            fin = torch.isfinite(c0)
            assert torch.equal(fin, torch.isfinite(c1)), 'masks differ'
            print(f'   cond q={q}: max err {float((c0[fin] - c1[fin]).abs().max()):.3e}')
        wf.set_inference_precision('fp64')
        if n == 20 and heads == 4:
            big = idx.repeat(60)[: 1 << 18].contiguous()
            for prec in ('fp64', 'tf32'):
                wf.log_psi_kernel(big, precision=prec); torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(3): wf.log_psi_kernel(big, precision=prec)
                torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
                print(f'   {prec}: {big.shape[0] / dt:.3e} amplitudes/s')
