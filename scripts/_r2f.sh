#!/bin/bash
# experiment: L2 cache hints on the slot reads (evict-first, no L1 allocation) and the filter words (evict-last) of the bit-sliced fused
# kernel (run under gpurun).  libanqs_b200_exp{0,1,2}.so were one-off builds of k1_fused_bs.cu with the three variants described in
# profiles/r2b_cache_hint_experiment.txt; the winner (evict-first slot reads for tables beyond the L2) is now in the shipped kernel.
mkdir -p gpurun_out
python - <<'P' 2>&1 | tee gpurun_out/cache_hint_experiment.txt
import os, sys, tempfile, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib
dev = torch.device('cuda:0'); main = _lib.lib()
name = 'anqs_local_energy_sample_aware_variant'
libs = {'shipped': main}
for m in (0, 1, 2):
    L = ctypes.CDLL(os.path.abspath(f'anqs_quantum_chemistry_b200/libanqs_b200_exp{m}.so'))
    getattr(L, name).restype = getattr(main, name).restype
    getattr(L, name).argtypes = getattr(main, name).argtypes
    libs[f'exp{m}'] = L
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
rows = 1 << 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    for keys in (1 << 23, 1 << 20):
        samples = synthetic.random_physical_samples(56, 7, 7, keys, seed=1)
        amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
        s = torch.from_numpy(samples.view(np.int64)).to(dev); a = torch.from_numpy(amps).to(dev)
        table = SampleTable(s, a)
        eloc = torch.empty(rows, dtype=torch.complex128, device=dev)
        sp = _lib.stream_ptr(dev)
        def go(L):
            _lib.check(getattr(L, name)(ham.tables, _lib.dptr(s), _lib.dptr(torch.view_as_real(a)), s.shape[0], 0, rows,
                                        _lib.dptr(table.slots), table.capacity, 7, 7, _lib.dptr(torch.view_as_real(eloc)), 2, sp))
        for L in libs.values(): go(L)
        torch.cuda.synchronize()
        for rnd in range(2):
            for tag, L in libs.items():
                ts = []
                for _ in range(4):
                    flush.fill_(1); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); go(L); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                print(keys, 'keys', tag, [round(t, 2) for t in ts], 'E sum', complex(eloc.sum()))
        del table, s, a
P
