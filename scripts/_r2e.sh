#!/bin/bash
# experiment: L2 access-policy variants for the presence filter of an 8.4M-key table (run under gpurun).  The modes are read
# from ANQS_EXP_L2 by a build with scripts/exp_l2_policy.patch applied to csrc/k1_fused_bs.cu (not part of the shipped library):
# 0 no window, 1 persisting + streaming misses (shipped), 2 persisting + normal misses, 3 hit ratio 0.6 + streaming, 4 0.6 + normal
mkdir -p gpurun_out
python - <<'P' 2>&1 | tee gpurun_out/l2_policy_experiment.txt
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib
dev = torch.device('cuda:0'); lib = _lib.lib()
keys, rows = 1 << 23, 1 << 20
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
samples = synthetic.random_physical_samples(56, 7, 7, keys, seed=1)
amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    s = torch.from_numpy(samples.view(np.int64)).to(dev); a = torch.from_numpy(amps).to(dev)
    table = SampleTable(s, a)
    eloc = torch.empty(rows, dtype=torch.complex128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sp = _lib.stream_ptr(dev)
    def go():
        _lib.check(lib.anqs_local_energy_sample_aware_variant(ham.tables, _lib.dptr(s), _lib.dptr(torch.view_as_real(a)), s.shape[0], 0, rows,
                                                              _lib.dptr(table.slots), table.capacity, 7, 7, _lib.dptr(torch.view_as_real(eloc)), 2, sp))
    go(); torch.cuda.synchronize()
    for mode in ('1', '0', '2', '3', '4', '1'):
        os.environ['ANQS_EXP_L2'] = mode
        ts = []
        for _ in range(5):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('mode', mode, [round(t, 2) for t in ts], 'E sum', complex(eloc.sum()))
P
