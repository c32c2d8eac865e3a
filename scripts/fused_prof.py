"""A few launches of the fused sample-aware local-energy kernel on the C5 shape (for ncu)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device('cuda:0')
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
samples = synthetic.random_physical_samples(56, 7, 7, rows, seed=1)
amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    s = torch.from_numpy(samples.view(np.int64)).to(dev).view(-1, 1)
    a = torch.from_numpy(amps).to(dev)
    table = SampleTable(s.view(-1), a)
    for _ in range(3):
        e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham', alpha_num=7, beta_num=7, table=table)[0]
    torch.cuda.synchronize()
    print('E sum', complex(e.sum()))
