"""Quick GPU check of the tiled enumeration (k1_enum.cu) against the untiled kernels: bit-equality + timing."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, synthetic, _lib

dev = torch.device('cuda:0')

def tm(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def run(n, n_el, irreps, rows, timing=False, complex_w=False, off_sector=False):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=irreps, seed=0)
    if complex_w:
        w = w.astype(np.complex128) * np.exp(1j * 0.3)
    na = nb = n_el // 2
    samples = synthetic.random_physical_samples(n, na, nb, rows, seed=1)
    if off_sector:  # half of the samples carry (na + 1, nb - 1) electrons: x' can still land in the (na, nb) sector
        other = synthetic.random_physical_samples(n, na + 1, nb - 1, rows, seed=2)
        samples = np.unique(np.concatenate((samples[: rows // 2], other[: rows // 2])))
    with tempfile.TemporaryDirectory() as tmp:
        hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tmp, rng_seed=0)
        ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
        s = torch.from_numpy(samples.view(np.int64)).to(dev)
        print(f'n={n} U={ham.unq_xy_masks_num} T={ham.term_num} enum_tiles={ham.enum_tiles} rows={s.shape[0]} real={ham.weights_real}')
        me = 'real' if ham.weights_real else 'complex'
        a = ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=False)
        b = ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=True)
        c = ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=True, filter_variant=1)
        for k in ('counts', 'offsets', 'dest', 'xprime', 'xy_ptr'):
            assert torch.equal(a[k], b[k]), k
            assert torch.equal(a[k], c[k]), k
        err = float((a['H'] - b['H']).abs().max()) if a['H'].numel() else 0.0
        print('  bit-equal lists; max |H_tiled - H_untiled| =', err, ' M =', a['xprime'].shape[0])
        assert err < 1e-11
        if timing:
            lib, sp = _lib.lib(), _lib.stream_ptr(dev)
            nrows = s.shape[0]
            m = a['xprime'].shape[0]
            bitmap = torch.empty(nrows * ham.bitmap_row_words, dtype=torch.int32, device=dev)
            work = torch.empty((int(lib.anqs_k1_enum_workspace(ham.tables, nrows)) + 3) // 4, dtype=torch.int32, device=dev)
            counts, offsets = b['counts'], b['offsets']
            t_f = tm(lambda: _lib.check(lib.anqs_k1_enum_filter(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(counts), _lib.dptr(bitmap), _lib.dptr(work), sp)))
            t_fp = tm(lambda: _lib.check(lib.anqs_k1_enum_filter_variant(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(counts), _lib.dptr(bitmap), _lib.dptr(work), 1, sp)))
            t_f0 = tm(lambda: _lib.check(lib.anqs_k1_filter(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(counts), _lib.dptr(bitmap), sp)))
            _lib.check(lib.anqs_k1_enum_filter(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(counts), _lib.dptr(bitmap), _lib.dptr(work), sp))
            hptr = _lib.dptr(torch.view_as_real(b['H'])) if me == 'complex' else _lib.dptr(b['H'])
            hc = 2 if me == 'complex' else 1
            t_e = tm(lambda: _lib.check(lib.anqs_k1_enum_emit(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(bitmap), _lib.dptr(offsets), _lib.dptr(work),
                                                              _lib.dptr(b['dest']), _lib.dptr(b['xprime']), _lib.dptr(None), hptr, hc, sp)))
            t_e_noh = tm(lambda: _lib.check(lib.anqs_k1_enum_emit(ham.tables, _lib.dptr(s), nrows, na, nb, _lib.dptr(bitmap), _lib.dptr(offsets), _lib.dptr(work),
                                                                  _lib.dptr(b['dest']), _lib.dptr(b['xprime']), _lib.dptr(None), _lib.dptr(None), 0, sp)))
            t_e0 = tm(lambda: _lib.check(lib.anqs_k1_emit(ham.tables, _lib.dptr(s), nrows, _lib.dptr(bitmap), _lib.dptr(offsets),
                                                          _lib.dptr(b['dest']), _lib.dptr(b['xprime']), _lib.dptr(None), hptr, hc, sp)))
            by = (12 + 8 * hc) * m
            print(f'  filter bit-sliced {t_f:.3f} ms (product {t_fp:.3f}, untiled {t_f0:.3f}); emit tiled {t_e:.3f} ms (no H {t_e_noh:.3f}; untiled {t_e0:.3f}); '
                  f'{by / ((t_f + t_e) * 1e-3) / 1e9:.0f} GB/s filter+emit, {by / (t_e * 1e-3) / 1e9:.0f} GB/s emit only')

run(12, 4, 1, 200)
run(20, 14, 1, 3000)
run(20, 14, 1, 500, complex_w=True)
run(36, 12, 8, 2000)
run(20, 14, 1, 1000, off_sector=True)
run(12, 4, 1, 200, off_sector=True, complex_w=True)
run(56, 14, 8, 1000)
run(56, 14, 8, 16384, timing=True)
run(56, 14, 8, 65536, timing=True)
