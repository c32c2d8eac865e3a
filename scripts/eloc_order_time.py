"""Does the ORDER of the rows matter for the fused local-energy kernel?  Samples of the MADE wave function at the C5 shape in the
order the sampler returns them (tree order), sorted, and randomly permuted; and the bench's random configurations."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, ParticleNumberSymmetry, SpinHalfProjectionSymmetry,
                                         LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig, SampleTable, synthetic)
dev = torch.device('cuda:0')
xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
hs = HilbertSpace(qubit_num=56, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=14),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
wf.set_inference_precision('tf32')
idx, cnt = wf.sample_stats(10 ** 6, seed=3)
idx = idx.view(-1)
with torch.no_grad():
    amps = wf.amplitude(idx.view(-1, 1))

def t(s, a, label):
    table = SampleTable(s, a)
    for _ in range(2):
        e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham', alpha_num=7, beta_num=7, table=table)[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham', alpha_num=7, beta_num=7, table=table)[0]
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'{label:28s} rows {s.shape[0]:8d}  {ms:7.2f} ms  {ms * 1e6 / s.shape[0]:6.2f} ns/row  sum {complex(e.sum()):.6f}')

t(idx, amps, 'sampler (tree) order')
srt, perm = torch.sort(idx)
t(srt, amps[perm], 'sorted')
rp = torch.randperm(idx.shape[0], device=dev)
t(idx[rp], amps[rp], 'random permutation')
rnd = torch.from_numpy(synthetic.random_physical_samples(56, 7, 7, idx.shape[0], seed=1).view(np.int64)).to(dev)
t(rnd, torch.from_numpy(synthetic.random_amplitudes(rnd.shape[0], seed=2)).to(dev), 'bench: random configurations')
