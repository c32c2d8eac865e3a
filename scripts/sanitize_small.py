"""Small invocations of the round-1 kernels (tiled enumeration with both filters, bit-sliced and per-sample fused local energy,
Gumbel select) on awkward sizes - meant for `compute-sanitizer --tool memcheck` (closed on this pool, so it only ran plain)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import (HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic, _lib,
                                         ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig)
dev = torch.device('cuda:0')
lib = _lib.lib()
for n, n_el, irreps, rows, off in ((12, 4, 1, 70, False), (20, 14, 1, 333, True), (56, 14, 8, 257, False)):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=irreps, seed=0)
    na = nb = n_el // 2
    samples = synthetic.random_physical_samples(n, na, nb, rows, seed=1)
    if off:
        samples = np.unique(np.concatenate((samples[: rows // 2], synthetic.random_physical_samples(n, na + 1, nb - 1, rows, seed=2)[: rows // 2])))
    amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
    with tempfile.TemporaryDirectory() as tmp:
        hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tmp, rng_seed=0)
        ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
        s = torch.from_numpy(samples.view(np.int64)).to(dev)
        a = torch.from_numpy(amps).to(dev)
        for force in (0, 1):
            c = ham.connected_configurations(s, na, nb, matrix_elements='real', tiled=True, filter_variant=force)
        c0 = ham.connected_configurations(s, na, nb, matrix_elements='real', tiled=False)
        assert torch.equal(c['xprime'], c0['xprime'])
        table = SampleTable(s, a)
        for choice in (1, 2):
            e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                   alpha_num=na, beta_num=nb, table=table, kernel_variant=choice)[0]
        torch.cuda.synchronize()
        print('ok', n, samples.shape[0], c['xprime'].shape[0], complex(e.sum()))
n, n_el = 12, 4
hs = HilbertSpace(qubit_num=n, device=dev, parent_dir=tempfile.mkdtemp(), rng_seed=0)
masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                 SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
torch.manual_seed(0)
wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
idx, f = wf.sample_indices_gumbel(100)
torch.cuda.synchronize()
print('gumbel ok', idx.shape[0], float(f.sum()))
