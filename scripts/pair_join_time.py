"""Pair-join ('trie' / 'all_to_all') against enumerate-and-probe ('ham') for the sample-aware local energy: C3 shape (20 qubits,
dense H, U = 2 536) and C5 shape (56 qubits, U = 23 157) at several sampled-set sizes."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, synthetic

dev = torch.device('cuda:0')


def tm(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for n, ne, irreps, sizes in ((20, 14, 1, (1000, 10000, 14400)), (56, 14, 8, (1000, 10000, 65536))):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=irreps, seed=0)
    hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    for N in sizes:
        samples = synthetic.random_physical_samples(n, ne // 2, ne // 2, N, seed=1)
        s = torch.from_numpy(samples.view(np.int64)).to(dev).view(-1, 1)
        a = torch.from_numpy(synthetic.random_amplitudes(samples.shape[0], seed=2)).to(dev)
        f = lambda v: ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham', alpha_num=ne // 2,
                                                         beta_num=ne // 2, kernel_variant=v)[0]
        e0, e3 = f(0), f(3)
        err = float((e0 - e3).abs().max())
        t0, t3 = tm(lambda: f(0)), tm(lambda: f(3))
        print(f'n={n} U={ham.unq_xy_masks_num} N={samples.shape[0]}: enumerate-and-probe {t0:.3f} ms, pair-join {t3:.3f} ms ({t0 / t3:.2f}x), max|dE| = {err:.2e}', flush=True)
