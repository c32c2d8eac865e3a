"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, through oracle/ref_shim.py) on small synthetic cases.

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.make_golden            # all groups
    python -m oracle.make_golden ham        # only the Hamiltonian / local-energy group

The reference has no tests or fixtures of its own; these files are what pins the
oracle (oracle/anqs_oracle.c, oracle/anqs_numpy.py) and, through it, the CUDA kernels.
Every array stored is either an input we generated or an output the reference produced.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from anqs_quantum_chemistry_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _ham_case(name, qubit_num, particle_num, n_irreps, sample_count, seed, keep_terms=None, store_lists=True,
              methods=('ham', 'trie', 'all_to_all'), chunk_size=20000):
    xy, yz, w = synthetic.synthetic_hamiltonian(qubit_num, n_irreps=n_irreps, seed=seed)
    if keep_terms is not None and keep_terms < xy.shape[0]:
        rng = np.random.default_rng(seed + 100)
        keep = np.sort(rng.choice(np.arange(1, xy.shape[0]), size=keep_terms - 1, replace=False))
        keep = np.concatenate(([0], keep))
        xy, yz, w = xy[keep], yz[keep], w[keep]
    # shuffle term order so that "original order inside each XY group" (PO:194-197) is exercised
    rng = np.random.default_rng(seed + 7)
    perm = rng.permutation(xy.shape[0])
    xy, yz, w = xy[perm], yz[perm], w[perm]
    terms = synthetic.pauli_arrays_to_terms(xy, yz, w, qubit_num)
    assert len(terms) == xy.shape[0]

    tmp = tempfile.mkdtemp(prefix='anqs_golden_')
    try:
        o = ref_shim.build_reference_objects(terms, qubit_num, particle_num, tmp)
        ham, hs = o.ham, o.hs
        na = nb = particle_num // 2
        if sample_count is None:
            samples = synthetic.all_physical_samples(qubit_num, na, nb)
        else:
            samples = synthetic.random_physical_samples(qubit_num, na, nb, sample_count, seed=seed + 1)
        amps = synthetic.random_amplitudes(samples.shape[0], seed=seed + 2)
        s_t = _t(samples.view(np.int64)).reshape(-1, 1)
        a_t = _t(amps)

        out = dict(qubit_num=qubit_num, particle_num=particle_num,
                   in_xy=xy.view(np.int64), in_yz=yz.view(np.int64), in_w=w,
                   samples=samples.view(np.int64), amps=amps,
                   unq_xy_masks=ham.unq_xy_masks.numpy().reshape(-1),
                   unq_xy_masks_inv=ham.unq_xy_masks_inv.numpy(),
                   unq_xy_to_yz_num=ham.unq_xy_to_yz_num.numpy(),
                   unq_xy_to_yz_start=ham.unq_xy_to_yz_start.numpy(),
                   rearranged_yz=ham.rearranged_yz.numpy().reshape(-1),
                   rearranged_weights=ham.rearranged_weights.numpy())

        # PO:527-567 candidates + filter on the whole batch as one chunk
        n = samples.shape[0]
        ptrs = torch.arange(n)
        dest, xp, xyptr, _ = ham.compute_candidates_for_coupling_via_ham(chunk_as_unq_batch_ptrs=ptrs,
                                                                           unq_batch_as_base_indices=s_t)
        dest, xp, xyptr, _ = ham.filter_candidates_for_coupling_via_ham(dest_as_chunk_ptrs=dest, src_as_base_indices=xp,
                                                                          coupling_xy_as_unq_ham_xy_ptrs=xyptr,
                                                                          alpha_num=na, beta_num=nb)
        H, _, _ = ham.compute_matrix_elements(x_primes=xp, ham_xy_pointers=xyptr)
        in_mask, in_ptr = hs.find_a_in_b(a=xp, b=s_t)
        out['conn_count_per_sample'] = np.bincount(dest.numpy(), minlength=n)
        if store_lists:
            out.update(conn_dest=dest.numpy(), conn_xprime=xp.numpy().reshape(-1), conn_xy_ptr=xyptr.numpy(),
                       conn_H=H.numpy(), conn_in_mask=in_mask.numpy(), conn_in_ptr=in_ptr.numpy())
        else:
            out.update(conn_H_sum=np.array(H.sum().item()), conn_xprime_xor=np.bitwise_xor.reduce(xp.numpy().reshape(-1)),
                       conn_in_count=np.array(int(in_mask.sum())))
        for method in methods:
            e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s_t, unq_batch_as_amps=a_t,
                                                         coupling_method=method, chunk_size=chunk_size,
                                                         alpha_num=na, beta_num=nb)
            out[f'eloc_{method}'] = e.numpy()
        # chunked evaluation must agree with single-chunk (PO:416-418)
        e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s_t, unq_batch_as_amps=a_t,
                                                     coupling_method='ham', chunk_size=max(1, n // 3),
                                                     alpha_num=na, beta_num=nb)
        out['eloc_ham_chunked'] = e.numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}.npz'), **out)
        print(f'{name}: T={xy.shape[0]} U={ham.unq_xy_masks_num} N={n} M={int(dest.shape[0])}')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _hilbert_case():
    """HS:121-284 helpers on random inputs, including negative words (bit 63 set)."""
    tmp = tempfile.mkdtemp(prefix='anqs_golden_')
    try:
        ref = ref_shim.load_reference()
        hs = ref.HilbertSpace(qubit_num=64, device=torch.device('cpu'), parent_dir=tmp, rng_seed=0,
                              popcount_mode='memory_efficient')
        rng = np.random.default_rng(11)
        a = rng.integers(-2**63, 2**63 - 1, size=257, dtype=np.int64)
        a[:5] = [0, -1, 1, -2**63, 2**63 - 1]
        dup = np.concatenate((a, a[::3], rng.integers(-50, 50, size=100, dtype=np.int64)))
        b = rng.permutation(np.unique(np.concatenate((a[::2], rng.integers(-2**63, 2**63 - 1, size=100, dtype=np.int64)))))
        pc = hs.popcount(_t(a).reshape(-1, 1)).numpy()
        srt, perm = hs.sort_base_idx(_t(dup).reshape(-1, 1))
        unq, inv = hs.compute_unique_indices(_t(dup).reshape(-1, 1))
        m, p = hs.find_a_in_b(a=_t(a).reshape(-1, 1), b=_t(b).reshape(-1, 1))
        hs20 = ref.HilbertSpace(qubit_num=20, device=torch.device('cpu'), parent_dir=tmp, rng_seed=0,
                                popcount_mode='memory_efficient')
        idx20 = rng.integers(0, 2**20, size=64, dtype=np.int64)
        vec20 = hs20.base_idx2base_vec(_t(idx20).reshape(-1, 1))
        back20 = hs20.base_vec2base_idx(vec20)
        vec64 = hs.base_idx2base_vec(_t(a[:32]).reshape(-1, 1))
        back64 = hs.base_vec2base_idx(vec64)
        np.savez_compressed(os.path.join(GOLDEN_DIR, 'hilbert.npz'), a=a, dup=dup, b=b, popcount=pc,
                            sorted=srt.numpy().reshape(-1), sort_perm=perm.numpy(),
                            unq=unq.numpy().reshape(-1), unq_inv=inv.numpy(), a_in_b_mask=m.numpy(), a_in_b_ptr=p.numpy(),
                            idx20=idx20, vec20=vec20.numpy(), back20=back20.numpy().reshape(-1),
                            vec64=vec64.numpy(), back64=back64.numpy().reshape(-1))
        print('hilbert: ok')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_ham():
    _ham_case('ham_n8_dense', 8, 4, 1, None, seed=0)
    _ham_case('ham_n12_dense', 12, 4, 1, 120, seed=0)
    _ham_case('ham_n14_dense', 14, 10, 1, None, seed=3, store_lists=False)
    _ham_case('ham_n20_dense', 20, 14, 1, 400, seed=0, store_lists=False, chunk_size=200)
    _ham_case('ham_n56_sparse', 56, 14, 8, 300, seed=0, keep_terms=3000, store_lists=False, chunk_size=100)
    _ham_case('ham_n64_sparse', 64, 16, 8, 200, seed=5, keep_terms=1500, store_lists=True, chunk_size=100)
    _hilbert_case()


GROUPS = {'ham': make_ham}


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    names = argv or list(GROUPS)
    for g in names:
        GROUPS[g]()


if __name__ == '__main__':
    main(sys.argv[1:])
