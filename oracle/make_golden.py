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


def make_ham_c5():
    # the headline Hamiltonian itself (bench.py: 56 qubits, 14 e-, 8 irreps, seed 0, all T = 114 305 terms): the reference's
    # own sample-aware local energies of 256 sampled configurations, coupling 'ham' only (its trie / all_to_all passes over
    # this table take tens of minutes on these cores and are covered by the smaller cases)
    _ham_case('ham_c5_full', 56, 14, 8, 256, seed=0, keep_terms=None, store_lists=False, methods=('ham',), chunk_size=64)
    # the inputs are regenerable (synthetic_hamiltonian(56, 8, seed 0) shuffled by default_rng(7)): keep checksums of them
    # and of the reference's tables instead of 2.6 MB of arrays
    path = os.path.join(GOLDEN_DIR, 'ham_c5_full.npz')
    g = dict(np.load(path))
    keep = {k: g[k] for k in ('qubit_num', 'particle_num', 'samples', 'amps', 'eloc_ham', 'eloc_ham_chunked', 'conn_count_per_sample',
                              'conn_H_sum', 'conn_xprime_xor', 'conn_in_count')}
    keep['in_checksums'] = np.array([int(np.bitwise_xor.reduce(g['in_xy'])), int(np.bitwise_xor.reduce(g['in_yz'])), int(g['in_xy'].shape[0])],
                                    dtype=np.int64)
    keep['in_w_sum'] = np.array([g['in_w'].sum().real, (np.abs(g['in_w']) ** 2).sum()])
    keep['tables_checksums'] = np.array([int(np.bitwise_xor.reduce(g['unq_xy_masks'])), int(g['unq_xy_masks'].shape[0]),
                                         int(np.bitwise_xor.reduce(g['rearranged_yz'])), int(g['unq_xy_to_yz_num'].sum())], dtype=np.int64)
    keep['rearranged_weights_probe'] = g['rearranged_weights'][::997]
    np.savez_compressed(path, **keep)


# ---- wave function, masks, samplers -------------------------------------------------------------------------------
def made_weights(n, Q, DM, depth=2, width=64, seed=0):
    """Deterministic (numpy PCG64) network weights shared by the golden generator and the tests, so that fixtures
    only need to store the seed.  Layout = reference parameter order: per sub-network, per layer, (weight, bias)."""
    rng = np.random.default_rng(seed)
    dims = [n] + [width] * depth + [Q * DM]
    nets = []
    for _net in range(2):
        layers = []
        for l in range(depth + 1):
            bound = 1.0 / np.sqrt(dims[l])
            layers.append((rng.uniform(-bound, bound, size=(dims[l + 1], dims[l])), rng.uniform(-bound, bound, size=dims[l + 1])))
        nets.append(layers)
    return nets


def load_weights_into(wf, nets):
    sd = {}
    for name, layers in zip(('log_abs_subnet', 'phase_subnet'), nets):
        for l, (w, b) in enumerate(layers):
            sd[f'{name}.layers.{l}.weight'] = torch.from_numpy(w.copy())
            sd[f'{name}.layers.{l}.bias'] = torch.from_numpy(b.copy())
    wf.load_state_dict(sd)


class _RintBinomial:
    """Stand-in for torch.distributions.Binomial in the reference run: sample() = rint(n p) (what draw_mode 0 of
    anqs_sampler_split_level and oracle/anqs_numpy.split_counts_rint compute)."""

    def __init__(self, total_count=None, probs=None):
        self.n, self.p = total_count.to(torch.float64), probs

    def sample(self):
        return torch.minimum(self.n, torch.clamp(torch.round(self.n * self.p), min=0.0))


def _anqs_case(name, n, n_el, sample_count, stats_num, gumbel_num, seed, z2=(), masking_depth=0):
    tmp = tempfile.mkdtemp(prefix='anqs_golden_')
    try:
        o = ref_shim.build_reference_objects(None, n, n_el, tmp, z2=z2, masking_depth=masking_depth)
        wf, masker, qg = o.wf, o.masker, o.wf.qubit_grouping
        Q, DM = qg.qudit_num, int(max(qg.qudit_dims))
        out = dict(qubit_num=n, particle_num=n_el, weight_seed=seed, qudit_num=Q, max_qudit_dim=DM, masking_depth=masking_depth,
                   z2_values=np.array([v for v, _ in z2], dtype=np.int64),
                   z2_masks=np.array([sum(1 << int(p) for p in pos) for _, pos in z2], dtype=np.int64))
        # initial weights under pt.manual_seed(0) (build_reference_objects seeds before constructing the ansatz)
        out['init_checksums'] = np.array([[float(p.sum()), float((p * p).sum()), float(p.reshape(-1)[0]), float(p.reshape(-1)[-1])]
                                          for p in wf.parameters()])
        out['param_num'] = wf.param_num
        nets = made_weights(n, Q, DM, seed=seed)
        load_weights_into(wf, nets)
        # tables
        out['memo'] = np.packbits(masker.memo.numpy())
        out['cont_mask_words'] = np.stack([(qg.qudit_idx2cont_mask_mul_table[q].numpy().astype(np.uint64)
                                            << np.arange(int(qg.qudit_dims[q]), dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
                                           for q in range(Q)])
        out['next_memo_masked_sum'] = np.array([int((qg.qudit_idx2memo_idx_mul_table[q] * qg.qudit_idx2cont_mask_mul_table[q]).sum())
                                                for q in range(Q)])
        na = nb = n_el // 2
        phys = synthetic.random_physical_samples(n, na, nb, sample_count * (8 if z2 else 1), seed=seed + 1)
        for v, pos in z2:  # keep the configurations every Z2 generator accepts
            zmask = np.uint64(sum(1 << int(p) for p in pos))
            par = np.array([bin(int(x & zmask)).count('1') & 1 for x in phys])
            phys = phys[(1 - 2 * par) == v]
        phys = phys[:sample_count]
        rng = np.random.default_rng(seed + 2)
        unphys = rng.integers(0, 2 ** min(n, 62), size=8, dtype=np.int64).astype(np.uint64)
        samples = np.concatenate((phys, unphys))
        s_t = _t(samples.view(np.int64)).reshape(-1, 1)
        base_vec = wf.base_idx2base_vec(s_t)
        with torch.no_grad():
            lp = wf.log_psi(base_vec)
            amp = wf.amplitude(s_t)
        out.update(samples=samples.view(np.int64), n_phys=phys.shape[0], log_psi=lp.numpy(), amplitude=amp.numpy())
        # conditional log-amplitudes of two levels for the prefixes of the physical samples
        for q in sorted({0, Q // 2, Q - 1}):
            with torch.no_grad():
                prefix_vec = base_vec[:phys.shape[0], :qg.qudit_starts[q]]
                rolling = masker.compute_rolling_acc_eigs(prefix_vec)[-1]
                mask = qg.qudit_idx2cont_mask_mul_table[q][masker.acc_eigs2memo_idx(rolling)]
                if wf.local_sampling_pattern[q] == 'DU':  # the mask the reference's own callers pass for such a qudit (ANQS:417-418)
                    mask = torch.ones_like(mask)
                full = torch.zeros((mask.shape[0], DM), dtype=torch.bool)
                full[:, :mask.shape[1]] = mask
                out[f'cond_log_abs_q{q}'] = wf.cond_log_abs(qudit_idx=q, base_vec=prefix_vec, mask=full).numpy()
        # gradient of sum_b Re(conj(c_b) log psi_b) with fixed random complex c, through the reference's autograd
        c = rng.standard_normal(phys.shape[0]) + 1j * rng.standard_normal(phys.shape[0])
        wf.zero_grad()
        lp = wf.log_psi(base_vec[:phys.shape[0]])
        loss = (torch.conj(_t(c)) * lp).real.sum()
        loss.backward()
        grad = wf.cat_grad.numpy()
        proj = np.random.default_rng(seed + 3).standard_normal((16, grad.shape[0]))
        out.update(grad_coeff=c, grad_loss=float(loss), grad_proj=proj @ grad, grad_norms=np.array([float(p.grad.norm()) for p in wf.parameters()]),
                   grad_head=grad[:64].copy(), grad_tail=grad[-64:].copy())
        # amplitude path: d/dtheta of sum_b Re(conj(c_b) psi_b) (exp on top of log psi, ANQS:483-485)
        wf.zero_grad()
        loss2 = (torch.conj(_t(c)) * wf.amplitude(s_t[:phys.shape[0]])).real.sum()
        loss2.backward()
        out['grad_amp_proj'] = proj @ wf.cat_grad.numpy()
        # count-splitting sampler with the binomial draw replaced by its rounded mean
        real_binomial = torch.distributions.Binomial
        torch.distributions.Binomial = _RintBinomial
        try:
            idx, cnt = wf.sample_stats(stats_num)
        finally:
            torch.distributions.Binomial = real_binomial
        out.update(stats_num=stats_num, stats_idx=idx.numpy().reshape(-1), stats_counts=cnt.numpy().real)
        # Gumbel top-k with uniforms from a recorded numpy stream
        urng = np.random.default_rng(seed + 4)
        real_rand = torch.rand
        torch.rand = lambda shape, **kw: torch.from_numpy(urng.random(tuple(shape)))
        try:
            gidx, gfreq = wf.sample_indices_gumbel(gumbel_num)
        finally:
            torch.rand = real_rand
        out.update(gumbel_num=gumbel_num, gumbel_idx=gidx.numpy().reshape(-1), gumbel_freqs=gfreq.numpy())
        np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}.npz'), **out)
        print(f'{name}: P={wf.param_num} stats_unique={idx.shape[0]} gumbel_unique={gidx.shape[0]}')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _nade_case(name, n, n_el, sample_count, stats_num, gumbel_num, seed):
    """NADE mode (the reference's default de_mode): one MLP pair per qudit.  Weights are the reference's own initial weights
    under pt.manual_seed(seed) - the drop-in reproduces them by constructing its modules in the same order."""
    tmp = tempfile.mkdtemp(prefix='anqs_golden_')
    try:
        o = ref_shim.build_reference_objects(None, n, n_el, tmp, de_mode='NADE', rng_seed=seed)
        wf, masker, qg = o.wf, o.masker, o.wf.qubit_grouping
        Q, DM = qg.qudit_num, int(max(qg.qudit_dims))
        out = dict(qubit_num=n, particle_num=n_el, seed=seed, qudit_num=Q, max_qudit_dim=DM, param_num=wf.param_num,
                   param_names=np.array([k for k, _ in wf.named_parameters()]),
                   init_checksums=np.array([[float(p.sum()), float((p * p).sum())] for p in wf.parameters()]))
        na = nb = n_el // 2
        phys = synthetic.random_physical_samples(n, na, nb, sample_count * (8 if z2 else 1), seed=seed + 1)
        for v, pos in z2:  # keep the configurations every Z2 generator accepts
            zmask = np.uint64(sum(1 << int(p) for p in pos))
            par = np.array([bin(int(x & zmask)).count('1') & 1 for x in phys])
            phys = phys[(1 - 2 * par) == v]
        phys = phys[:sample_count]
        rng = np.random.default_rng(seed + 2)
        samples = np.concatenate((phys, rng.integers(0, 2 ** min(n, 62), size=8, dtype=np.int64).astype(np.uint64)))
        s_t = _t(samples.view(np.int64)).reshape(-1, 1)
        base_vec = wf.base_idx2base_vec(s_t)
        with torch.no_grad():
            lp = wf.log_psi(base_vec)
            amp = wf.amplitude(s_t)
        out.update(samples=samples.view(np.int64), n_phys=phys.shape[0], log_psi=lp.numpy(), amplitude=amp.numpy())
        for q in sorted({0, Q // 2, Q - 1}):
            with torch.no_grad():
                prefix_vec = base_vec[:phys.shape[0], :qg.qudit_starts[q]]
                rolling = masker.compute_rolling_acc_eigs(prefix_vec)[-1]
                mask = qg.qudit_idx2cont_mask_mul_table[q][masker.acc_eigs2memo_idx(rolling)]
                out[f'cond_log_abs_q{q}'] = wf.cond_log_abs(qudit_idx=q, base_vec=prefix_vec, mask=mask).numpy()
        c = rng.standard_normal(phys.shape[0]) + 1j * rng.standard_normal(phys.shape[0])
        wf.zero_grad()
        loss = (torch.conj(_t(c)) * wf.log_psi(base_vec[:phys.shape[0]])).real.sum()
        loss.backward()
        grad = wf.cat_grad.numpy()
        proj = np.random.default_rng(seed + 3).standard_normal((16, grad.shape[0]))
        out.update(grad_coeff=c, grad_loss=float(loss), grad_proj=proj @ grad, grad_norms=np.array([float(p.grad.norm()) for p in wf.parameters()]))
        real_binomial = torch.distributions.Binomial
        torch.distributions.Binomial = _RintBinomial
        try:
            idx, cnt = wf.sample_stats(stats_num)
        finally:
            torch.distributions.Binomial = real_binomial
        out.update(stats_num=stats_num, stats_idx=idx.numpy().reshape(-1), stats_counts=cnt.numpy().real)
        urng = np.random.default_rng(seed + 4)
        real_rand = torch.rand
        torch.rand = lambda shape, **kw: torch.from_numpy(urng.random(tuple(shape)))
        try:
            gidx, gfreq = wf.sample_indices_gumbel(gumbel_num)
        finally:
            torch.rand = real_rand
        out.update(gumbel_num=gumbel_num, gumbel_idx=gidx.numpy().reshape(-1), gumbel_freqs=gfreq.numpy())
        np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}.npz'), **out)
        print(f'{name}: P={wf.param_num} stats_unique={idx.shape[0]} gumbel_unique={gidx.shape[0]}')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_nade():
    _nade_case('nade_n12', 12, 4, 100, 10 ** 4, 64, seed=0)
    _nade_case('nade_n20', 20, 14, 200, 10 ** 6, 300, seed=1)
    _nade_case('nade_n56', 56, 14, 200, 3000, 200, seed=2)


def make_anqs():
    _anqs_case('anqs_n12', 12, 4, 100, 10 ** 4, 64, seed=0)
    _anqs_case('anqs_n20', 20, 14, 200, 10 ** 6, 300, seed=1)
    _anqs_case('anqs_n56', 56, 14, 200, 3000, 200, seed=2)
    _anqs_case('anqs_n14', 14, 10, 100, 10 ** 5, 500, seed=3)
    # the reference's default masker level ('z2': Z2 generators on top of N and S_z) and a non-zero masking depth
    _anqs_case('anqs_z2_n12', 12, 4, 100, 10 ** 4, 40, seed=4, z2=((1, (0, 3, 5, 8)), (-1, (1, 2, 10))))
    _anqs_case('anqs_md1_n20', 20, 14, 100, 10 ** 5, 100, seed=5, masking_depth=1)


# ---- one whole VMC iteration (EXP:626-679 without SR): sample -> amplitudes -> local energies -> loss -> backward ----
def _vmc_case(name, n, n_el, n_irreps, sample_num, seed):
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=n_irreps, seed=seed)
    terms = synthetic.pauli_arrays_to_terms(xy, yz, w, n)
    tmp = tempfile.mkdtemp(prefix='anqs_golden_')
    try:
        o = ref_shim.build_reference_objects(terms, n, n_el, tmp)
        ref, wf, ham = o.ref, o.wf, o.ham
        Q, DM = wf.qubit_grouping.qudit_num, int(max(wf.qubit_grouping.qudit_dims))
        load_weights_into(wf, made_weights(n, Q, DM, seed=seed))
        out = dict(qubit_num=n, particle_num=n_el, n_irreps=n_irreps, ham_seed=seed, weight_seed=seed, sample_num=sample_num)
        # sampling through calculations.sample (SMP:51-101), count splitting with the deterministic binomial stand-in
        real_binomial = torch.distributions.Binomial
        torch.distributions.Binomial = _RintBinomial
        try:
            res, _, _, _ = ref.sample(wf=wf, config=ref.SamplingConfig(sample_indices=False, sample_num=sample_num))
        finally:
            torch.distributions.Binomial = real_binomial
        indices, perm = wf.sort_base_idx(res.indices)                       # EXP:511
        counts = res.counts[perm]
        out.update(indices=indices.numpy().reshape(-1), counts=counts.numpy().real)
        # compute_loss (EXP:548-611) with loss_type='sample_aware_e_loc', coupling 'ham'
        wf.zero_grad()
        amps = wf.amplitude(indices)
        sr = ref.SamplingResult(indices=indices, counts=counts)
        cfg = ref.LocalEnergyCalculationConfig(use_tree_for_candidates='ham')
        le, _ = ref.compute_local_energies(wf=wf, sampling_result=sr, sampled_amps=amps.detach(), ham=ham, config=cfg, sample_aware=True)
        est = le.sample_aware_e_loc_mc_est
        loss = 2 * (est.freqs * torch.log(torch.conj(amps)) * (est.values - est.mean)).sum().real   # EXP:609
        loss.backward()
        grad = wf.cat_grad.numpy()
        proj = np.random.default_rng(seed + 3).standard_normal((16, grad.shape[0]))
        out.update(amps=amps.detach().numpy(), eloc=est.values.numpy(), energy_mean=complex(est.mean), energy_var=complex(est.var),
                   loss=float(loss), grad_proj=proj @ grad, grad_norm=float(np.linalg.norm(grad)), grad_head=grad[:64].copy(),
                   grad_tail=grad[-64:].copy())
        # gradient post-processing (PG:55-70): stochastic reconfiguration on the top-25 samples (SR:88-136) + clipping
        jac = wf.compute_cat_log_jac(indices[:8])
        jproj = np.random.default_rng(seed + 5).standard_normal((grad.shape[0], 4))
        out.update(log_jac_proj=(jac.detach().numpy() @ jproj), log_jac_head=jac.detach().numpy()[:, :32].copy())
        ref.process_grad(wf=wf, sampling_result=sr, config=ref.ProcessGradConfig())
        g_sr = wf.cat_grad.numpy()
        out.update(sr_grad_proj=proj @ g_sr, sr_grad_norm=float(np.linalg.norm(g_sr)), sr_grad_head=g_sr[:64].copy())
        # the same without regularisation (pseudo-inverse branch) on a fresh gradient
        wf.cat_grad = torch.from_numpy(grad.copy())
        ref.process_grad(wf=wf, sampling_result=sr, config=ref.ProcessGradConfig(sr_config=ref.SRConfig(use_reg=False, max_indices_num=10),
                                                                                clip_grad_norm=False, renorm_grad=True))
        g_sr2 = wf.cat_grad.numpy()
        out.update(sr2_grad_proj=proj @ g_sr2, sr2_grad_norm=float(np.linalg.norm(g_sr2)))
        # the full (not sample-aware) local energy through the old code path (CLE:117-163 -> PO:326-393, 992-1105)
        cfg_old = ref.LocalEnergyCalculationConfig(use_tree_for_candidates='ham', code_version='old')
        le_full, metrics = ref.compute_local_energies(wf=wf, sampling_result=sr, sampled_amps=amps.detach(), ham=ham, config=cfg_old,
                                                      sample_aware=False)
        out.update(eloc_full=le_full.full_e_loc_mc_est.values.numpy(), eloc_full_aware=le_full.sample_aware_e_loc_mc_est.values.numpy(),
                   full_energy_mean=complex(le_full.full_e_loc_mc_est.mean),
                   non_sampled_unq=int(metrics.non_sampled_unq_x_primes_num), candidates=int(metrics.candidate_x_primes_num))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}.npz'), **out)
        print(f'{name}: N_unq={indices.shape[0]} <E>={complex(est.mean):.6f} full <E>={complex(le_full.full_e_loc_mc_est.mean):.6f} '
              f'loss={float(loss):.6e} non-sampled unique x\'={int(metrics.non_sampled_unq_x_primes_num)}')
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_vmc():
    _vmc_case('vmc_n12', 12, 4, 1, 300, seed=0)          # C1 shape (LiH STO-3G): 12 qubits, 4 electrons
    _vmc_case('vmc_n20', 20, 14, 1, 20000, seed=1)       # C3 shape (N2 STO-3G): 20 qubits, 14 electrons, T = 14 251


# ---- transformer network (BASELINE config 3): logits of the reference's own TransformerMADE module -----------------------------
def _tfm_case(name, n, n_el, depth, head_num, seed, sample_count=64):
    ref_shim.load_reference()
    from nqs.stochastic.ansatzes.legacy.anqs_primitives.made.transformer_made import TransformerMADE
    torch.manual_seed(seed)
    net = TransformerMADE(dim=64, out_dim=4, depth=depth, qubit_num=n, head_num=head_num, dtype=torch.float64)
    na = nb = n_el // 2
    phys = synthetic.random_physical_samples(n, na, nb, sample_count, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    samples = np.concatenate((phys, rng.integers(0, 2 ** n, size=8, dtype=np.int64).astype(np.uint64)))
    bits = torch.from_numpy(((samples[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int64))
    net.eval()
    with torch.no_grad():
        logits = net(bits)                                  # [B, n + 1, 4]
        short = net(bits[:, :n // 2])                      # a prefix: what the samplers evaluate at level n // 2
    out = dict(qubit_num=n, particle_num=n_el, depth=depth, head_num=head_num, seed=seed, samples=samples.view(np.int64),
               n_phys=phys.shape[0], logits=logits.numpy(), prefix_logits=short.numpy(),
               param_names=np.array([k for k, _ in net.named_parameters()]),
               param_checksums=np.array([[float(p.sum()), float((p * p).sum())] for p in net.parameters()]))
    np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}.npz'), **out)
    print(f'{name}: P={sum(p.numel() for p in net.parameters())} logits {tuple(logits.shape)}')
    # gradient golden: autograd through the reference's module of  sum_i a_i log|psi(x_i)| + b_i arg psi(x_i)  over the physical
    # samples, log psi = masked, normalised sum of the chosen conditionals (ANQS:392-405; masks from the numpy oracle)
    from oracle import anqs_numpy as onp
    masks = onp.NumberSpinMasks(n, n_el, qubit_per_qudit=1)
    xs = phys.view(np.uint64)
    allowed = np.stack([masks.cont_mask[t][masks.memo_idx_of_prefix(xs, t)] for t in range(n)], axis=1)       # [B, n, 2]
    a, b = rng.standard_normal(phys.shape[0]), rng.standard_normal(phys.shape[0])
    net.zero_grad()
    o = net(bits[:phys.shape[0]])[:, :n, :].reshape(phys.shape[0], n, 2, 2)
    re = torch.where(torch.from_numpy(allowed), o[..., 0], torch.full_like(o[..., 0], -np.inf))
    re = re - 0.5 * torch.logsumexp(2.0 * re, dim=-1, keepdim=True)
    pick = bits[:phys.shape[0]].unsqueeze(-1)
    log_abs = torch.gather(re, -1, pick).squeeze(-1).sum(-1)
    phase = torch.gather(o[..., 1], -1, pick).squeeze(-1).sum(-1)
    (torch.from_numpy(a) * log_abs + torch.from_numpy(b) * phase).sum().backward()
    grads = {f'grad_{i:02d}': p.grad.numpy() for i, (_, p) in enumerate(net.named_parameters())}
    np.savez_compressed(os.path.join(GOLDEN_DIR, f'{name}_grad.npz'), a=a, b=b, log_abs=log_abs.detach().numpy(),
                        phase=phase.detach().numpy(), **grads)
    print(f'{name}_grad: |g| = {float(sum((g * g).sum() for g in grads.values())) ** 0.5:.6g}')


def make_tfm():
    _tfm_case('tfm_n20', 20, 14, depth=2, head_num=4, seed=7)     # C3 shape
    _tfm_case('tfm_n12', 12, 4, depth=1, head_num=8, seed=8)
    _tfm_case('tfm_n14', 14, 10, depth=3, head_num=2, seed=9)     # head_dim 32: the wide-head path of the kernel


GROUPS = {'ham': make_ham, 'ham_c5': make_ham_c5, 'anqs': make_anqs, 'nade': make_nade, 'vmc': make_vmc, 'tfm': make_tfm}


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    names = argv or list(GROUPS)
    for g in names:
        GROUPS[g]()


if __name__ == '__main__':
    main(sys.argv[1:])
