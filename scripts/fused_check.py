"""GPU check: bit-sliced fused local-energy kernel vs the warp-per-sample kernel (bit-level agreement of E_loc) + timing."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib

dev = torch.device('cuda:0')
lib = _lib.lib()

def tm(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def run(n, n_el, irreps, rows, timing=False, complex_w=False, off_sector=False, clustered=False):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=irreps, seed=0)
    if complex_w:
        w = w.astype(np.complex128) * np.exp(0.3j)
    na = nb = n_el // 2
    if clustered:
        samples = synthetic.clustered_physical_samples(n, na, nb, rows, seed=7, mean_rank=2.0)
    else:
        samples = synthetic.random_physical_samples(n, na, nb, rows, seed=1)
    if off_sector:
        other = synthetic.random_physical_samples(n, na + 1, nb - 1, rows, seed=2)
        samples = np.unique(np.concatenate((samples[: rows // 2], other[: rows // 2])))
    amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
    with tempfile.TemporaryDirectory() as tmp:
        hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tmp, rng_seed=0)
        ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
        s = torch.from_numpy(samples.view(np.int64)).to(dev).view(-1, 1)
        a = torch.from_numpy(amps).to(dev)
        table = SampleTable(s.view(-1), a)
        def go(variant):
            return ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                                      alpha_num=na, beta_num=nb, table=table, kernel_variant=variant)[0]
        e_new = go(2)
        e_old = go(1)
        err = float((e_new - e_old).abs().max()); scale = float(e_old.abs().max())
        print(f'n={n} rows={s.shape[0]} complex={complex_w} off_sector={off_sector} clustered={clustered}: max|dE| = {err:.3e} (scale {scale:.3e})')
        assert err <= 1e-11 * max(1.0, scale)
        if timing:
            t_old = tm(lambda: go(1))
            t_new = tm(lambda: go(2))
            print(f'  per-sample kernel {t_old:.3f} ms, bit-sliced {t_new:.3f} ms  ({s.shape[0] / t_new / 1e3:.3e} E_loc/s)')

run(12, 4, 1, 200)
run(20, 14, 1, 3000)
run(20, 14, 1, 6000, clustered=True)
run(20, 14, 1, 500, complex_w=True)
run(20, 14, 1, 1000, off_sector=True)
run(36, 12, 8, 2000)
run(56, 14, 8, 1000)
run(56, 14, 8, 4000, clustered=True)
run(56, 14, 8, 10000, timing=True)
run(56, 14, 8, 65536, timing=True)
run(56, 14, 8, 1 << 20, timing=True)
