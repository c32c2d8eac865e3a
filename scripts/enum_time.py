"""Timing of the tiled enumeration on the C5 shape: python scripts/enum_time.py [rows ...] (env ANQS_ENUM_BAND_ROWS = rows per band)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, synthetic, _lib

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def tm(fn, reps=8, warm=2, cold=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    lib, sp = _lib.lib(), _lib.stream_ptr(dev)
    for rows in [int(a) for a in sys.argv[1:]] or [16384]:
        samples = synthetic.random_physical_samples(56, 7, 7, rows, seed=1)
        s = torch.from_numpy(samples.view(np.int64)).to(dev)
        b = ham.connected_configurations(s, 7, 7, with_xy_ptr=False, matrix_elements='real')
        m = b['xprime'].shape[0]
        bitmap = torch.empty(rows * ham.bitmap_row_words, dtype=torch.int32, device=dev)
        work = torch.empty((int(lib.anqs_k1_enum_workspace(ham.tables, rows)) + 3) // 4, dtype=torch.int32, device=dev)
        counts, offsets = b['counts'], b['offsets']
        f = lambda: _lib.check(lib.anqs_k1_enum_filter(ham.tables, _lib.dptr(s), rows, 7, 7, _lib.dptr(counts), _lib.dptr(bitmap), _lib.dptr(work), sp))
        e = lambda hc: _lib.check(lib.anqs_k1_enum_emit(ham.tables, _lib.dptr(s), rows, 7, 7, _lib.dptr(bitmap), _lib.dptr(offsets), _lib.dptr(work),
                                                        _lib.dptr(b['dest']), _lib.dptr(b['xprime']), _lib.dptr(None), _lib.dptr(b['H']) if hc else _lib.dptr(None), hc, sp))
        f()
        for cold in (False, True):
            t_f, t_e, t_n = tm(f, cold=cold), tm(lambda: e(1), cold=cold), tm(lambda: e(0), cold=cold)
            print(f'band={os.environ.get("ANQS_ENUM_BAND_ROWS", "default")} rows={rows} M={m} {"L2 flushed" if cold else "back to back"}: filter {t_f:.3f} emit {t_e:.3f} (no H {t_n:.3f}) ms; '
                  f'{20 * m / (t_f + t_e) / 1e6:.0f} GB/s filter+emit = {20 * m / (t_f + t_e) / 1e6 / 6547.5:.3f} of HBM peak; emit only {20 * m / t_e / 1e6:.0f} GB/s', flush=True)
        del b, bitmap, work
