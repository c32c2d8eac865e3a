"""Dense 56-qubit Hamiltonian (1 irrep: U = 184 591 masks, T ~ 8.8e5 terms): the big-table paths of the enumeration (many tiles,
several filter chunks) and of the fused local energy (many product tiles) against the untiled / per-sample kernels."""
import sys, os, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib
dev = torch.device('cuda:0'); lib = _lib.lib()
t0 = time.time()
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=1, seed=0)
print(f'generated T={xy.shape[0]} in {time.time() - t0:.1f} s', flush=True)
na = nb = 7
samples = synthetic.random_physical_samples(56, na, nb, 40000, seed=1)
amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    t0 = time.time()
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    ham.tables
    print(f'tables U={ham.unq_xy_masks_num} T={ham.term_num} enum_tiles={ham.enum_tiles} built in {time.time() - t0:.1f} s', flush=True)
    s = torch.from_numpy(samples.view(np.int64)).to(dev)
    a = torch.from_numpy(amps).to(dev)
    rows = s[:512].contiguous()
    ref = ham.connected_configurations(rows, na, nb, matrix_elements='real', tiled=False)
    out = ham.connected_configurations(rows, na, nb, matrix_elements='real', tiled=True)
    for k in ('counts', 'offsets', 'dest', 'xprime', 'xy_ptr'):
        assert torch.equal(ref[k], out[k]), k
    err = float((ref['H'] - out['H']).abs().max())
    print(f'enumeration: {ref["xprime"].shape[0]} connections of 512 rows identical; max |dH| = {err:.2e}', flush=True)
    assert err < 1e-11 * max(1.0, float(ref['H'].abs().max()))
    def tm(fn, reps=3):
        fn(); torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
    big = s[:4096].contiguous()
    m = int(ham.connected_configurations(big, na, nb, with_dest=False, with_xy_ptr=False)['xprime'].shape[0])
    t_t = tm(lambda: ham.connected_configurations(big, na, nb, with_xy_ptr=False, matrix_elements='real', tiled=True))
    t_u = tm(lambda: ham.connected_configurations(big, na, nb, with_xy_ptr=False, matrix_elements='real', tiled=False))
    print(f'enumeration of 4096 rows ({m} connections, incl. allocation and scan): tiled {t_t:.2f} ms = {20 * m / t_t / 1e6:.0f} GB/s, untiled {t_u:.2f} ms', flush=True)
    table = SampleTable(s, a)
    res = {}
    for choice in (1, 2):
        f = lambda choice=choice: ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                                     alpha_num=na, beta_num=nb, table=table, kernel_variant=choice)[0]
        res[choice] = f()
        print(('per-sample' if choice == 1 else 'bit-sliced'), f'fused kernel, {s.shape[0]} rows: {tm(f):.2f} ms', flush=True)
    err = float((res[1] - res[2]).abs().max())
    print(f'fused kernels agree to {err:.2e} (scale {float(res[1].abs().max()):.2e})')
    assert err < 1e-11 * max(1.0, float(res[1].abs().max()))
