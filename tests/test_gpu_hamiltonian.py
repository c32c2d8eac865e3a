"""GPU parity: the CUDA local-energy path (through the C ABI) against the reference's golden outputs and,
on larger seeded inputs, against the CPU oracle.  Bit-exact for connected configurations, pointers and
counts; 1e-10 for matrix elements and local energies (fp64)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, HAM_CASES, HAM_CASES_WITH_LISTS
from anqs_quantum_chemistry_b200 import HilbertSpace, PauliObservable, PauliArraysOperator, SampleTable, synthetic
from oracle import hamiltonian_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _ham(g, tmp_path):
    n = int(g['qubit_num'])
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    return hs, PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(g['in_xy'], g['in_yz'], g['in_w'], n))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('case', HAM_CASES)
def test_connected_list_matches_reference(case, tmp_path):
    g = load_golden(case)
    hs, ham = _ham(g, tmp_path)
    na = nb = int(g['particle_num']) // 2
    conn = ham.connected_configurations(_dev(g['samples']), na, nb, matrix_elements='complex')
    np.testing.assert_array_equal(conn['counts'].cpu().numpy(), g['conn_count_per_sample'])
    off = conn['offsets'].cpu().numpy()
    np.testing.assert_array_equal(off, np.concatenate(([0], np.cumsum(g['conn_count_per_sample']))))
    xp = conn['xprime'].cpu().numpy()
    H = conn['H'].cpu().numpy()
    mask, ptr = hs.find_a_in_b(a=conn['xprime'].view(-1, 1), b=_dev(g['samples']).view(-1, 1))
    if case in HAM_CASES_WITH_LISTS:
        np.testing.assert_array_equal(conn['dest'].cpu().numpy(), g['conn_dest'])
        np.testing.assert_array_equal(xp, g['conn_xprime'])
        np.testing.assert_array_equal(conn['xy_ptr'].cpu().numpy(), g['conn_xy_ptr'])
        assert np.abs(H - g['conn_H']).max() < 1e-10
        np.testing.assert_array_equal(mask.cpu().numpy(), g['conn_in_mask'])
        np.testing.assert_array_equal(ptr.cpu().numpy(), g['conn_in_ptr'])
    else:
        assert np.bitwise_xor.reduce(xp) == g['conn_xprime_xor']
        assert abs(H.sum() - g['conn_H_sum']) < 1e-9
        assert int(mask.sum().item()) == int(g['conn_in_count'])
    # real-valued matrix elements (20-byte rows) agree with the complex ones
    if ham.weights_real:
        conn_r = ham.connected_configurations(_dev(g['samples']), na, nb, with_dest=False, with_xy_ptr=False, matrix_elements='real')
        np.testing.assert_array_equal(conn_r['H'].cpu().numpy(), H.real)
        np.testing.assert_array_equal(conn_r['xprime'].cpu().numpy(), xp)


@pytest.mark.parametrize('case', HAM_CASES)
def test_local_energy_matches_reference(case, tmp_path):
    g = load_golden(case)
    hs, ham = _ham(g, tmp_path)
    na = nb = int(g['particle_num']) // 2
    s, a = _dev(g['samples']).view(-1, 1), _dev(g['amps'])
    scale = max(1.0, np.abs(g['eloc_ham']).max())
    for method in ('ham', 'trie', 'all_to_all'):
        e, e2, metrics = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a,
                                                            coupling_method=method, chunk_size=20000, alpha_num=na, beta_num=nb)
        assert e.dtype == torch.complex128 and tuple(e.shape) == (s.shape[0],)
        assert np.abs(e.cpu().numpy() - g[f'eloc_{method}']).max() < 1e-10 * scale
    # the pair-join kernel (what 'trie' / 'all_to_all' run for large mask lists and small batches) against the reference's trie path
    e_pj = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='trie',
                                              alpha_num=na, beta_num=nb, kernel_variant=3)[0]
    assert np.abs(e_pj.cpu().numpy() - g['eloc_trie']).max() < 1e-10 * scale
    with pytest.raises(NotImplementedError):
        ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='hamming_ball',
                                           alpha_num=na, beta_num=nb)
    # rows walked in the strided order (what large batches do by default), whole batch and a window of it
    e_s = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                             alpha_num=na, beta_num=nb, row_order='strided')[0]
    assert np.abs(e_s.cpu().numpy() - g['eloc_ham']).max() < 1e-10 * scale
    if s.shape[0] >= 6:
        lo_s, ln_s = s.shape[0] // 3, s.shape[0] // 2
        w_s = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham', alpha_num=na,
                                                 beta_num=nb, row_start=lo_s, row_len=ln_s, row_order='strided')[0]
        assert np.abs(w_s.cpu().numpy() - g['eloc_ham'][lo_s:lo_s + ln_s]).max() < 1e-10 * scale
    # a window of rows (the multi-GPU shard path) equals the same rows of the full evaluation
    n = s.shape[0]
    lo, ln = n // 3, n // 2
    for variant in (0, 3):   # the fused kernel and the pair-join kernel
        e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                               alpha_num=na, beta_num=nb, kernel_variant=variant)[0]
        w, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                                     alpha_num=na, beta_num=nb, row_start=lo, row_len=ln, kernel_variant=variant)
        assert torch.equal(w, e[lo:lo + ln])


def test_headline_hamiltonian_matches_reference_output(tmp_path):
    """The bench's own Hamiltonian (56 qubits, T = 114 305, U = 23 157) against output of the REFERENCE on it (golden
    ham_c5_full: PauliObservable.compute_var_local_energy_proxy, coupling 'ham', 256 sampled configurations): tables by
    checksum, connection counts exactly, sample-aware local energies to 1e-10 through every kernel variant."""
    from conftest import c5_full_inputs
    g, xy, yz, w = c5_full_inputs()
    hs = HilbertSpace(qubit_num=56, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    tc = g['tables_checksums']
    assert ham.unq_xy_masks_num == int(tc[1]) and int(ham.unq_xy_to_yz_num.sum()) == int(tc[3])
    assert int(np.bitwise_xor.reduce(ham.unq_xy_masks.cpu().numpy().reshape(-1))) == int(tc[0])
    assert int(np.bitwise_xor.reduce(ham.rearranged_yz.cpu().numpy().reshape(-1))) == int(tc[2])
    np.testing.assert_array_equal(ham.rearranged_weights.cpu().numpy()[::997], g['rearranged_weights_probe'])
    s, a = _dev(g['samples']).view(-1, 1), _dev(g['amps'])
    conn = ham.connected_configurations(_dev(g['samples']), 7, 7, matrix_elements='complex')
    np.testing.assert_array_equal(np.bincount(conn['dest'].cpu().numpy(), minlength=s.shape[0]), g['conn_count_per_sample'])
    assert int(np.bitwise_xor.reduce(conn['xprime'].cpu().numpy().reshape(-1))) == int(g['conn_xprime_xor'])
    assert abs(complex(conn['H'].sum().item()) - complex(g['conn_H_sum'])) < 1e-8
    scale = max(1.0, np.abs(g['eloc_ham']).max())
    for variant in (0, 1, 2, 3):   # chosen by size, warp-per-sample, bit-sliced, pair-join
        e = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                               alpha_num=7, beta_num=7, kernel_variant=variant)[0]
        assert np.abs(e.cpu().numpy() - g['eloc_ham']).max() < 1e-10 * scale, variant


@pytest.mark.parametrize('case', HAM_CASES_WITH_LISTS)
def test_reference_method_surface(case, tmp_path):
    """find_sampled_and_coupled_via_ham (PO:569-600) and compute_matrix_elements (PO:255-324)."""
    g = load_golden(case)
    hs, ham = _ham(g, tmp_path)
    na = nb = int(g['particle_num']) // 2
    s = _dev(g['samples']).view(-1, 1)
    n = s.shape[0]
    dest, src_ptr, src_idx, xy_ptr, metrics, seconds = ham.find_sampled_and_coupled(
        chunk_as_unq_batch_ptrs=torch.arange(n, device=DEV), unq_batch_as_base_indices=s, coupling_method='ham',
        symmetric=False, alpha_num=na, beta_num=nb, metrics=None)
    m = g['conn_in_mask']
    np.testing.assert_array_equal(dest.cpu().numpy(), g['conn_dest'][m])
    np.testing.assert_array_equal(src_ptr.cpu().numpy(), g['conn_in_ptr'][m])
    np.testing.assert_array_equal(src_idx.cpu().numpy().reshape(-1), g['conn_xprime'][m])
    np.testing.assert_array_equal(xy_ptr.cpu().numpy(), g['conn_xy_ptr'][m])
    assert dest.dtype == torch.int64 and tuple(src_idx.shape) == (int(m.sum()), 1)
    assert metrics.candidate_x_primes_num == g['conn_xprime'].shape[0]
    H, yz_num, secs = ham.compute_matrix_elements(x_primes=_dev(g['conn_xprime']).view(-1, 1), ham_xy_pointers=_dev(g['conn_xy_ptr']))
    assert np.abs(H.cpu().numpy() - g['conn_H']).max() < 1e-10
    assert yz_num == int(g['unq_xy_to_yz_num'][g['conn_xy_ptr']].sum())


def test_hilbert_helpers_match_reference(tmp_path):
    g = load_golden('hilbert')
    hs = HilbertSpace(qubit_num=64, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    a = _dev(g['a']).view(-1, 1)
    np.testing.assert_array_equal(hs.popcount(a).cpu().numpy(), g['popcount'])
    odd = _dev(np.concatenate((g['a'], g['a'][:1])))[1:]  # unaligned start, odd length
    np.testing.assert_array_equal(hs.popcount(odd.view(-1, 1)).cpu().numpy(), g['popcount'][1:].tolist() + [g['popcount'][0]])
    inplace = a.clone()
    out = hs.popcount_(inplace)
    np.testing.assert_array_equal(out.cpu().numpy().reshape(-1), g['popcount'])
    s, p = hs.sort_base_idx(_dev(g['dup']).view(-1, 1))
    np.testing.assert_array_equal(s.cpu().numpy().reshape(-1), g['sorted'])
    np.testing.assert_array_equal(p.cpu().numpy(), g['sort_perm'])
    u, inv = hs.compute_unique_indices(_dev(g['dup']).view(-1, 1))
    np.testing.assert_array_equal(u.cpu().numpy().reshape(-1), g['unq'])
    np.testing.assert_array_equal(inv.cpu().numpy(), g['unq_inv'])
    m, ptr = hs.find_a_in_b(a=a, b=_dev(g['b']).view(-1, 1))
    np.testing.assert_array_equal(m.cpu().numpy(), g['a_in_b_mask'])
    np.testing.assert_array_equal(ptr.cpu().numpy(), g['a_in_b_ptr'])
    assert hs.popcount(torch.empty((0, 1), dtype=torch.int64, device=DEV)).numel() == 0


@pytest.mark.parametrize('qubits,electrons,irreps,n_samples', [(24, 10, 2, 3000), (56, 14, 8, 1500)])
def test_against_oracle_on_seeded_inputs(qubits, electrons, irreps, n_samples, tmp_path):
    """Sizes the CPU oracle finishes in seconds, including the headline 56-qubit shape (U ~ 2.3e4, streamed
    when the table does not fit in shared memory is covered by test_streamed_table)."""
    xy, yz, w = synthetic.synthetic_hamiltonian(qubits, n_irreps=irreps, seed=4)
    na = nb = electrons // 2
    samples = synthetic.random_physical_samples(qubits, na, nb, n_samples, seed=5)
    # make the sampled set connected: add single and double excitations of the first samples
    tab = orc.Tables(xy, yz, w)
    _, xp, _ = orc.candidates_ham(samples, 0, 4, tab, na, nb)
    samples = np.unique(np.concatenate((samples, xp.view(np.uint64)[:: max(1, xp.shape[0] // n_samples)])))
    amps = synthetic.random_amplitudes(samples.shape[0], seed=6)
    hs = HilbertSpace(qubit_num=qubits, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, qubits))
    s, a = _dev(samples.view(np.int64)), _dev(amps)
    e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                 alpha_num=na, beta_num=nb)
    e_ref = orc.local_energy_sample_aware(samples, amps, tab, na, nb)
    scale = max(1.0, np.abs(e_ref).max())
    assert np.abs(e.cpu().numpy() - e_ref).max() < 1e-10 * scale
    sub = 256
    conn = ham.connected_configurations(s[:sub], na, nb, matrix_elements='real')
    dest, xp, ptr = orc.candidates_ham(samples, 0, sub, tab, na, nb)
    np.testing.assert_array_equal(conn['dest'].cpu().numpy(), dest)
    np.testing.assert_array_equal(conn['xprime'].cpu().numpy(), xp)
    np.testing.assert_array_equal(conn['xy_ptr'].cpu().numpy(), ptr)
    H = orc.matrix_elements(xp, ptr, tab)
    assert np.abs(conn['H'].cpu().numpy() - H.real).max() < 1e-10 * max(1.0, np.abs(H).max())
    # hermiticity of the restricted Hamiltonian: sum_i conj(psi_i) (H psi)_i is real
    a_np = amps
    assert abs((np.conj(a_np) * (e.cpu().numpy() * a_np)).sum().imag) < 1e-9


def test_streamed_table(tmp_path):
    """U > 25600 forces the double-buffered TMA tile path (dense 36-qubit shape: U = 29 836)."""
    xy, yz, w = synthetic.synthetic_hamiltonian(36, n_irreps=1, seed=2)
    na = nb = 6
    samples = synthetic.random_physical_samples(36, na, nb, 300, seed=3)
    tab = orc.Tables(xy, yz, w)
    assert tab.unq_xy_masks_num > 25600
    _, xp, _ = orc.candidates_ham(samples, 0, 2, tab, na, nb)
    samples = np.unique(np.concatenate((samples, xp.view(np.uint64)[::7])))
    amps = synthetic.random_amplitudes(samples.shape[0], seed=6)
    hs = HilbertSpace(qubit_num=36, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 36))
    s, a = _dev(samples.view(np.int64)), _dev(amps)
    e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                 alpha_num=na, beta_num=nb)
    e_ref = orc.local_energy_sample_aware(samples, amps, tab, na, nb)
    assert np.abs(e.cpu().numpy() - e_ref).max() < 1e-10 * max(1.0, np.abs(e_ref).max())
    conn = ham.connected_configurations(s[:64], na, nb, matrix_elements='real')
    dest, xp, ptr = orc.candidates_ham(samples, 0, 64, tab, na, nb)
    np.testing.assert_array_equal(conn['xprime'].cpu().numpy(), xp)
    np.testing.assert_array_equal(conn['xy_ptr'].cpu().numpy(), ptr)
    np.testing.assert_array_equal(conn['dest'].cpu().numpy(), dest)


@pytest.mark.parametrize('qubits,electrons,irreps,n_samples', [(20, 14, 1, 6000), (56, 14, 8, 4000)])
def test_clustered_samples_and_filter_spread(qubits, electrons, irreps, n_samples, tmp_path):
    """Sample sets concentrated around the Hartree-Fock determinant: many samples share their alpha string (so the
    line-blocked presence filter is heavily loaded on a few lines) and a large share of the connected configurations
    is itself sampled (the matrix-element path is busy).  Every forced filter spread G = 0..6 and the data-driven
    choice must give the oracle's local energies."""
    xy, yz, w = synthetic.synthetic_hamiltonian(qubits, n_irreps=irreps, seed=4)
    na = nb = electrons // 2
    samples = synthetic.clustered_physical_samples(qubits, na, nb, n_samples, seed=7, mean_rank=2.0)
    amps = synthetic.random_amplitudes(samples.shape[0], seed=8)
    tab = orc.Tables(xy, yz, w)
    e_ref = orc.local_energy_sample_aware(samples, amps, tab, na, nb)
    scale = max(1.0, np.abs(e_ref).max())
    hs = HilbertSpace(qubit_num=qubits, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, qubits))
    s, a = _dev(samples.view(np.int64)), _dev(amps)
    # hits are frequent in this workload
    conn = ham.connected_configurations(s[:128], na, nb, with_dest=False, with_xy_ptr=False)
    mask, _ = hs.find_a_in_b(a=conn['xprime'].view(-1, 1), b=s.view(-1, 1))
    assert float(mask.double().mean()) > (0.05 if qubits == 20 else 0.001)
    for spread in (None, 0, 1, 3, 6):
        table = SampleTable(s, a, spread_bits=spread)
        e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                     alpha_num=na, beta_num=nb, table=table)
        assert np.abs(e.cpu().numpy() - e_ref).max() < 1e-10 * scale, spread
        g, over = table.filter_info()
        assert g == (spread if spread is not None else g) and 0 <= g <= 6
        assert all(over[i] >= over[i + 1] for i in range(6))


def test_filter_spread_is_chosen_from_the_data(tmp_path):
    """Uniform samples have ~1 key per alpha string -> G = 0; a set with one dominant alpha string -> G > 0."""
    uni = _dev(synthetic.random_physical_samples(56, 7, 7, 50000, seed=9).view(np.int64))
    g_uni, over = SampleTable(uni).filter_info()
    assert g_uni == 0 and over[0] == 0
    # all beta strings on top of a single alpha string: 20000 keys want the same line
    base = synthetic.random_physical_samples(56, 7, 7, 20000, seed=10)
    one_alpha = np.unique((base & np.uint64(0xAAAAAAAAAAAAAAAA)) | (base[0] & np.uint64(0x5555555555555555)))
    g_hot, over = SampleTable(_dev(one_alpha.view(np.int64))).filter_info()
    assert g_hot == 6 and over[0] == one_alpha.shape[0]


def test_large_table_with_duplicated_keys():
    """A table beyond the L2 cache (2^22 slots, 128 MiB) with duplicated keys: every occupied slot holds its key's LAST position
    and that position's amplitude (sequential scatter_ semantics, HS:263-284) - the insert kernel (k2_hash.cu) keeps four keys
    in flight per thread without fences and leaves racing amplitude stores of duplicates to its fix-up pass -, the all-ones
    key (the EMPTY sentinel) lives in its own slot, and a rebuild in place gives the same table (as a set of occupied slots:
    which slot a colliding key ends up in depends on the order the inserts arrive in)."""
    from anqs_quantum_chemistry_b200 import _lib
    n0, ndup = 1_250_000, 60_000
    base = synthetic.random_physical_samples(56, 7, 7, n0, seed=21).view(np.int64)
    keys = np.concatenate([base, base[1000:1000 + ndup], base[500:500 + ndup // 2], np.array([-1, -1], np.int64)])
    amps = synthetic.random_amplitudes(keys.shape[0], seed=22)
    s, a = _dev(keys), _dev(amps)
    table = SampleTable(s, a)
    cap = table.capacity
    assert cap == 1 << 22

    def occupied(t):
        rows = t.slots[:(cap + 1) * 4].view(cap + 1, 4)
        rows = rows[rows[:, 1] >= 0]                            # position >= 0 (the sentinel key's slot included)
        return rows[torch.argsort(rows[:, 0])]
    occ = occupied(table)
    assert occ.shape[0] == n0 + 1
    # positions: the last occurrence wins
    last = {}
    for j in range(n0, keys.shape[0]):
        last[int(keys[j])] = j
    expect = np.arange(keys.shape[0])
    for j in range(keys.shape[0]):
        if int(keys[j]) in last:
            expect[j] = last[int(keys[j])]
    ptr = torch.empty(keys.shape[0], dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib().anqs_hash_probe(_lib.dptr(table.slots), cap, _lib.dptr(s), keys.shape[0], _lib.dptr(ptr), _lib.dptr(None),
                                          _lib.stream_ptr(DEV)))
    assert np.array_equal(ptr.cpu().numpy(), expect)
    winners = occ[:, 1]
    assert torch.equal(torch.sort(winners).values, torch.from_numpy(np.unique(expect)).to(DEV))
    assert torch.equal(occ[:, 2:].contiguous().view(torch.float64), torch.view_as_real(a)[winners])
    for _ in range(3):
        table.rebuild(s, a)
        assert torch.equal(occupied(table), occ)
    # the bit-sliced fused kernel reads the slots of a table of this size (2^22 slots and more) with an evict-first L2 policy
    # (k1_fused_bs.cu:hash_lookup_stream): same local energies as the warp-per-sample kernel, which reads them plainly
    xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        hs = HilbertSpace(qubit_num=56, device=DEV, parent_dir=tmp, rng_seed=0)
        ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
        rows, out = 8192, {}
        for variant in (1, 2):
            e = torch.empty(rows, dtype=torch.complex128, device=DEV)
            _lib.check(_lib.lib().anqs_local_energy_sample_aware_variant(ham.tables, _lib.dptr(s), _lib.dptr(torch.view_as_real(a)), s.shape[0], 0, rows,
                                                                         _lib.dptr(table.slots), cap, 7, 7, _lib.dptr(torch.view_as_real(e)), variant,
                                                                         _lib.stream_ptr(DEV)))
            out[variant] = e.cpu().numpy()
        assert np.abs(out[1] - out[2]).max() < 1e-10 * max(1.0, np.abs(out[1]).max())
        assert np.abs(out[1]).max() > 0


def test_multi_tile_product_layout(tmp_path):
    """The headline Hamiltonian (56 qubits, 8 irreps: U = 23 157) does not fit one 200 KB tile: the fused kernel
    re-streams two tiles per group of samples.  Also rows past the end of a partial group (n not a multiple of 32)."""
    xy, yz, w = synthetic.synthetic_hamiltonian(56, n_irreps=8, seed=0)
    na = nb = 7
    samples = synthetic.clustered_physical_samples(56, na, nb, 3001, seed=3, mean_rank=2.5)
    amps = synthetic.random_amplitudes(samples.shape[0], seed=6)
    tab = orc.Tables(xy, yz, w)
    hs = HilbertSpace(qubit_num=56, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, 56))
    s, a = _dev(samples.view(np.int64)), _dev(amps)
    e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s.view(-1, 1), unq_batch_as_amps=a, coupling_method='ham',
                                                 alpha_num=na, beta_num=nb)
    e_ref = orc.local_energy_sample_aware(samples, amps, tab, na, nb)
    assert np.abs(e.cpu().numpy() - e_ref).max() < 1e-10 * max(1.0, np.abs(e_ref).max())


def test_edge_cases(tmp_path):
    g = load_golden('ham_n8_dense')
    hs, ham = _ham(g, tmp_path)
    empty = torch.empty(0, dtype=torch.int64, device=DEV)
    conn = ham.connected_configurations(empty, 2, 2, matrix_elements='complex')
    assert conn['xprime'].numel() == 0 and conn['offsets'].cpu().tolist() == [0]
    e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=empty.view(-1, 1),
                                                 unq_batch_as_amps=torch.empty(0, dtype=torch.complex128, device=DEV),
                                                 coupling_method='ham', alpha_num=2, beta_num=2)
    assert e.numel() == 0
    # a single sample: only the diagonal couples
    s = _dev(g['samples'][:1]).view(-1, 1)
    a = _dev(g['amps'][:1])
    e, _, _ = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham', alpha_num=2, beta_num=2)
    H = orc.dense_matrix_from_arrays(g['in_xy'], g['in_yz'], g['in_w'], 8)
    x = int(g['samples'][0])
    assert abs(e.cpu().numpy()[0] - H[x, x]) < 1e-12
    # an electron count nobody satisfies: no connections at all
    conn = ham.connected_configurations(_dev(g['samples']), 3, 1)
    assert conn['xprime'].numel() == 0 and int(conn['counts'].sum().item()) == 0


@pytest.mark.parametrize('qubits,electrons,irreps,rows,complex_w,off_sector', [
    (12, 4, 1, 200, False, False), (20, 14, 1, 1500, False, False), (20, 14, 1, 400, True, False),
    (20, 14, 1, 600, False, True), (12, 4, 1, 150, True, True), (36, 12, 8, 700, False, False), (56, 14, 8, 600, False, False),
    (36, 12, 1, 300, False, False), (36, 12, 1, 200, False, True)])   # dense 36 qubits: 933 bitmap words = two filter chunks
def test_tiled_enumeration_equals_untiled(qubits, electrons, irreps, rows, complex_w, off_sector, tmp_path):
    """The tile-resident kernels (bit-sliced or product-layout filter, pattern-table matrix elements; k1_enum.cu) and the
    untiled pair (k1_connected.cu) emit the same ordered list bit for bit, and matrix elements that agree to 1e-12 —
    also for complex weights and for samples outside the (N_alpha, N_beta) sector, where the pattern tables do not apply."""
    from anqs_quantum_chemistry_b200 import _lib
    xy, yz, w = synthetic.synthetic_hamiltonian(qubits, n_irreps=irreps, seed=5)
    if complex_w:
        w = w.astype(np.complex128) * np.exp(0.3j)
    na = nb = electrons // 2
    samples = synthetic.random_physical_samples(qubits, na, nb, rows, seed=11)
    if off_sector:
        other = synthetic.random_physical_samples(qubits, na + 1, nb - 1, rows, seed=12)
        samples = np.unique(np.concatenate((samples[: rows // 2], other[: rows // 2])))
    hs = HilbertSpace(qubit_num=qubits, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, qubits))
    assert ham.enum_tiles > 0
    s = _dev(samples.view(np.int64))
    me = 'complex' if complex_w else 'real'
    ref = ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=False)
    outs = [ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=True),
            ham.connected_configurations(s, na, nb, matrix_elements=me, tiled=True, filter_variant=1)]  # product-layout filter
    assert ref['xprime'].shape[0] > 0
    for out in outs:
        for k in ('counts', 'offsets', 'dest', 'xprime', 'xy_ptr'):
            assert torch.equal(ref[k], out[k]), k
        scale = max(1.0, float(ref['H'].abs().max()))
        assert float((ref['H'] - out['H']).abs().max()) < 1e-12 * scale
    # against the CPU oracle as well (first rows)
    sub = min(64, samples.shape[0])
    tab = orc.Tables(xy, yz, w)
    dest, xp, ptr = orc.candidates_ham(samples, 0, sub, tab, na, nb)
    m = xp.shape[0]
    np.testing.assert_array_equal(outs[0]['xprime'][:m].cpu().numpy(), xp)
    np.testing.assert_array_equal(outs[0]['xy_ptr'][:m].cpu().numpy(), ptr)
    H = orc.matrix_elements(xp, ptr, tab)
    got = outs[0]['H'][:m].cpu().numpy()
    assert np.abs(got - (H if complex_w else H.real)).max() < 1e-10 * max(1.0, np.abs(H).max())


@pytest.mark.parametrize('qubits,electrons,irreps,rows,complex_w,off_sector,clustered', [
    (12, 4, 1, 200, False, False, False), (20, 14, 1, 3000, False, False, False), (20, 14, 1, 5000, False, False, True),
    (20, 14, 1, 500, True, False, False), (20, 14, 1, 1000, False, True, False), (36, 12, 8, 1500, False, False, False),
    (56, 14, 8, 3000, False, False, True), (56, 14, 8, 40000, False, False, False), (36, 12, 1, 1200, False, False, False)])
def test_bit_sliced_local_energy_equals_per_sample_kernel(qubits, electrons, irreps, rows, complex_w, off_sector, clustered, tmp_path):
    """The two fused sample-aware kernels (a warp per group of 32 samples with bit-sliced electron-count tests, k1_fused_bs.cu;
    a warp per sample, k1_fused.cu) and the CPU oracle give the same local energies to 1e-10, also for complex weights, for
    clustered sample sets (many hits, loaded filter lines), samples outside the sector, and when several warps share a group."""
    from anqs_quantum_chemistry_b200 import _lib
    xy, yz, w = synthetic.synthetic_hamiltonian(qubits, n_irreps=irreps, seed=3)
    if complex_w:
        w = w.astype(np.complex128) * np.exp(0.3j)
    na = nb = electrons // 2
    if clustered:
        samples = synthetic.clustered_physical_samples(qubits, na, nb, rows, seed=7, mean_rank=2.0)
    else:
        samples = synthetic.random_physical_samples(qubits, na, nb, rows, seed=13)
    if off_sector:
        other = synthetic.random_physical_samples(qubits, na + 1, nb - 1, rows, seed=14)
        samples = np.unique(np.concatenate((samples[: rows // 2], other[: rows // 2])))
    amps = synthetic.random_amplitudes(samples.shape[0], seed=15)
    hs = HilbertSpace(qubit_num=qubits, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, qubits))
    s, a = _dev(samples.view(np.int64)).view(-1, 1), _dev(amps)
    table = SampleTable(s.view(-1), a)
    res = {}
    for choice in (1, 2, 0):
        res[choice] = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                                         alpha_num=na, beta_num=nb, table=table, kernel_variant=choice)[0].cpu().numpy()
    scale = max(1.0, np.abs(res[1]).max())
    assert np.abs(res[2] - res[1]).max() < 1e-11 * scale
    assert np.abs(res[0] - res[1]).max() < 1e-11 * scale
    # the bit-sliced kernel adds a sample's contributions in a fixed order (private accumulators per warp, lane-ordered sums
    # inside a batch of resolved hits, warps added in order): the same call gives the same bits every time
    for _ in range(3):
        again = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='ham',
                                                   alpha_num=na, beta_num=nb, table=table, kernel_variant=2)[0].cpu().numpy()
        assert np.array_equal(again, res[2])
    if samples.shape[0] <= 40000:  # the pair-join kernel (what 'trie' / 'all_to_all' run on small batches): same energies
        e_pj = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='trie',
                                                  alpha_num=na, beta_num=nb, kernel_variant=3)[0].cpu().numpy()
        assert np.abs(e_pj - res[1]).max() < 1e-11 * scale
        lo, ln = samples.shape[0] // 4, samples.shape[0] // 2   # a row window (the multi-GPU shard path)
        e_win = ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=a, coupling_method='trie', alpha_num=na,
                                                   beta_num=nb, kernel_variant=3, row_start=lo, row_len=ln)[0].cpu().numpy()
        assert np.array_equal(e_win, e_pj[lo:lo + ln])
    if samples.shape[0] <= 6000:
        e_ref = orc.local_energy_sample_aware(samples, amps, orc.Tables(xy, yz, w), na, nb)
        assert np.abs(res[2] - e_ref).max() < 1e-10 * scale


def test_full_local_energy_on_a_fresh_observable(tmp_path):
    """compute_local_energies(sample_aware=False) as the first call on a new PauliObservable (the device tables and their
    properties are created on demand) agrees with the dense-matrix local energy H psi / psi over the whole sector."""
    from anqs_quantum_chemistry_b200 import ParticleNumberSymmetry, SpinHalfProjectionSymmetry, LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig
    n, n_el = 12, 4
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=1, seed=0)
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=(ParticleNumberSymmetry(hilbert_space=hs, particle_num=n_el),
                                                                     SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0)))
    torch.manual_seed(0)
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ANQSConfig(de_mode='MADE'))
    sector = synthetic.random_physical_samples(n, n_el // 2, n_el // 2, 10 ** 4, seed=1)  # all 225 states of the sector
    assert sector.shape[0] == 225
    sub = sector[::3].copy()
    with torch.no_grad():
        psi = wf.amplitude(_dev(sector.view(np.int64)).view(-1, 1)).cpu().numpy()
        a = wf.amplitude(_dev(sub.view(np.int64)).view(-1, 1))
        full, aware, metrics = ham.compute_local_energies(wf=wf, sampled_indices=_dev(sub.view(np.int64)).view(-1, 1), sampled_amps=a,
                                                          sample_aware=False)
    tab = orc.Tables(xy, yz, w)
    e_all = orc.local_energy_sample_aware(sector, psi, tab, n_el // 2, n_el // 2)  # every connected state is in the set: exact H psi / psi
    pos = np.searchsorted(sector, sub)
    assert np.abs(full.cpu().numpy() - e_all[pos]).max() < 1e-10 * max(1.0, np.abs(e_all).max())
    assert metrics.non_sampled_unq_x_primes_num > 0


def test_full_size_properties(tmp_path):
    """BASELINE.json's full size (C5: 56 qubits, T = 114 305, 2^20 unique samples), where no CPU oracle finishes: properties that
    do not depend on the size.  Local energies: scale invariance E_loc[c psi] = E_loc[psi], hermiticity of the restricted
    Hamiltonian (sum conj(psi) H psi real), a row window equals the slice of the whole, the two fused kernels agree.
    Enumeration (65 536 rows, 2.5e8 connections): every x' has the sector's electron counts, x' ^ x[dest] is exactly the mask
    xy_ptr names, xy_ptr rises strictly inside every sample, counts / offsets are consistent, and the three filters agree."""
    from anqs_quantum_chemistry_b200 import _lib
    n, na, nb = 56, 7, 7
    xy, yz, w = synthetic.synthetic_hamiltonian(n, n_irreps=8, seed=0)
    samples = synthetic.random_physical_samples(n, na, nb, 1 << 20, seed=1)
    amps = synthetic.random_amplitudes(samples.shape[0], seed=2)
    hs = HilbertSpace(qubit_num=n, device=DEV, parent_dir=str(tmp_path), rng_seed=0)
    ham = PauliObservable(hilbert_space=hs, of_qubit_operator=PauliArraysOperator(xy, yz, w, n))
    assert ham.term_num == 114305 and ham.unq_xy_masks_num == 23157
    s, a = _dev(samples.view(np.int64)).view(-1, 1), _dev(amps)
    assert s.shape[0] == 1 << 20

    def eloc(amps_t, **kw):
        return ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=s, unq_batch_as_amps=amps_t, coupling_method='ham',
                                                  alpha_num=na, beta_num=nb, **kw)[0]

    e = eloc(a)
    scale = float(e.abs().max())
    assert torch.isfinite(torch.view_as_real(e)).all()
    # homogeneous of degree zero in psi
    e_c = eloc(a * (0.37 - 1.2j))
    assert float((e_c - e).abs().max()) < 1e-11 * max(1.0, scale)
    # <psi|H|psi> restricted to the sampled set is real
    hpsi = (a.conj() * (e * a)).sum()
    assert abs(float(hpsi.imag)) < 1e-9 * max(1.0, abs(float(hpsi.real)))
    # a row window is the slice of the whole (same table)
    table = SampleTable(s.view(-1), a)
    lo, ln = 123457, 300001
    e_win = eloc(a, table=table, row_start=lo, row_len=ln)
    assert torch.equal(e_win, e[lo:lo + ln])
    # the warp-per-sample kernel agrees with the bit-sliced one at this size
    e_ps = eloc(a, table=table, kernel_variant=1)
    assert float((e_ps - e).abs().max()) < 1e-11 * max(1.0, scale)
    del e_c, e_win, e_ps, table

    rows = s.view(-1)[: 1 << 16].contiguous()
    conn = ham.connected_configurations(rows, na, nb, matrix_elements='real')
    m = conn['xprime'].shape[0]
    assert m == int(conn['counts'].sum()) == int(conn['offsets'][-1]) and m > 2.4e8
    assert torch.equal(conn['offsets'][1:] - conn['offsets'][:-1], conn['counts'])
    xp, dest, ptr = conn['xprime'], conn['dest'].long(), conn['xy_ptr'].long()
    even = torch.tensor(0x5555555555555555, dtype=torch.int64, device=DEV)
    assert bool((hs.popcount(xp & even) == na).all()) and bool((hs.popcount(xp & ~even) == nb).all())
    assert torch.equal(xp ^ rows[dest], ham.unq_xy_masks.to(DEV).view(-1)[ptr])
    same = dest[1:] == dest[:-1]
    assert bool((dest[1:] >= dest[:-1]).all()) and bool((ptr[1:][same] > ptr[:-1][same]).all())
    assert torch.isfinite(conn['H']).all()
    counts_ref = conn['counts'].clone()
    del conn, xp, dest, ptr, same
    for tiled, force in ((True, 1), (False, 0)):   # product-layout filter, flat popcount filter
        other = ham.connected_configurations(rows[:8192], na, nb, with_dest=False, with_xy_ptr=False, tiled=tiled, filter_variant=force)
        assert torch.equal(other['counts'], counts_ref[:8192])
