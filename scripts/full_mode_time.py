"""Throughput of the FULL (not sample-aware) local energy, PO:992-1105: enumeration with matrix elements -> join against the
sampled set -> unique of the non-sampled x' -> wf.amplitude on them -> accumulate.  20 qubits (C3 shape) and 56 qubits (C5)."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry,
                                         SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, synthetic)
dev = torch.device('cuda:0')
for n, n_el, irreps, rows, tf32 in ((20, 14, 1, 2000, False), (56, 14, 8, 4096, False), (56, 14, 8, 4096, True)):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=irreps, seed=0)
    hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(prefix='anqs_full_'), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(0)
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
    if tf32:
        wf.set_inference_precision('tf32')
    samples = synthetic.random_physical_samples(n, n_el // 2, n_el // 2, rows, seed=1)
    s = torch.from_numpy(samples.view(np.int64)).to(dev).view(-1, 1)
    with torch.no_grad():
        a = wf.amplitude(s)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            full, aware, m = ham.compute_local_energies(wf=wf, sampled_indices=s, sampled_amps=a, sample_aware=False, chunk_size=2048)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'n={n} rows={s.shape[0]} nn={"tf32" if tf32 else "f64"}: {dt * 1e3:.1f} ms -> {s.shape[0] / dt:.3e} full E_loc/s; '
          f'x\' {m.sampled_x_primes_num + m.non_sampled_x_primes_num} ({m.non_sampled_unq_x_primes_num} unique non-sampled), '
          f'amplitude time {m.eval_non_sampled_amps_time * 1e3:.1f} ms')
