"""One filter + one emit launch of the tiled enumeration on the C5 shape (for ncu)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device('cuda:0')
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
samples = synthetic.random_physical_samples(56, 7, 7, rows, seed=1)
with tempfile.TemporaryDirectory() as tmp:
    hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tmp, rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    s = torch.from_numpy(samples.view(np.int64)).to(dev)
    for _ in range(2):
        c = ham.connected_configurations(s, 7, 7, with_xy_ptr=False, matrix_elements='real')
    torch.cuda.synchronize()
    print('M =', c['xprime'].shape[0])
