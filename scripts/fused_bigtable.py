"""Fused local-energy kernel on a 1M-row window of an 8M-key table (the per-rank work of the 8-GPU weak-scaling bench)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib
dev = torch.device('cuda:0')
lib = _lib.lib()
keys = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
rows = 1 << 20
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
samples = synthetic.random_physical_samples(56, 7, 7, keys, seed=1)
amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    s = torch.from_numpy(samples.view(np.int64)).to(dev)
    a = torch.from_numpy(amps).to(dev)
    table = SampleTable(s, a)
    eloc = torch.empty(rows, dtype=torch.complex128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sp = _lib.stream_ptr(dev)
    def go():
        _lib.check(lib.anqs_local_energy_sample_aware_variant(ham.tables, _lib.dptr(s), _lib.dptr(torch.view_as_real(a)), s.shape[0], 0, rows,
                                                              _lib.dptr(table.slots), table.capacity, 7, 7, _lib.dptr(torch.view_as_real(eloc)), choice, sp))
    for choice in (2, 1):
        ts = []
        for _ in range(6):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(('bit-sliced' if choice == 2 else 'per-sample'), f'{keys} keys, {rows} rows:', [round(t, 2) for t in ts], 'E sum', complex(eloc.sum()))
