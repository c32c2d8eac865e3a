"""PauliObservable drop-in (reference: nqs/nqs/stochastic/observables/pauli_observable.py:89-1105).

Same constructor keywords, attributes, table tensors, `.npy` cache file names and method signatures as the
reference, so `experiments/calculations/compute_local_energies.py:75-163` can drive it unchanged.  All
compute goes through the sm_100a kernels of libanqs_b200.so:

  compute_var_local_energy_proxy  -> anqs_local_energy_sample_aware (fused filter+probe+matrix element+sum)
  find_sampled_and_coupled_via_ham -> anqs_k1_filter / anqs_k1_emit + anqs_hash_probe
  compute_matrix_elements         -> anqs_matrix_elements
  compute_local_energies (full)   -> anqs_k1_emit with matrix elements, hash join, wf.amplitude on the
                                     de-duplicated non-sampled configurations, anqs_accumulate_rows

The three working coupling methods of the reference ('ham', 'all_to_all', 'trie') produce the same numbers
(SURVEY.md §4).  'ham' enumerates the U connected configurations of every sample and probes the sampled set (fused
kernel); 'trie' / 'all_to_all' join the sampled configurations pairwise (k1_pairs.cu: N^2 XOR + POPC tests, survivors looked
up among the unique masks) while the batch is small enough for that to be the cheaper algorithm, and fall back to the fused
kernel beyond.  'hamming_ball' is broken in the reference (PO:698) and is rejected.
"""
import os
import time
from typing import Tuple

import math
import numpy as np
import torch as pt

from . import _lib
from .abstract_hilbert_space_object import AbstractHilbertSpaceObject
from .hilbert_space import SampleTable


class LocalEnergyMetrics:
    """Same field names as the reference's Config subclass (PO:25-86)."""
    FIELDS = ('candidate_x_primes_num', 'sampled_unq_x_primes_num', 'sampled_x_primes_num', 'sampled_coupled_yz_num',
              'non_sampled_unq_x_primes_num', 'non_sampled_x_primes_num', 'non_sampled_coupled_yz_num',
              'candidates_time', 'filter_candidates_time', 'find_a_in_b_time', 'find_sampled_and_coupled_time',
              'sampled_matrix_elements_time', 'sampled_scatter_time', 'non_sampled_matrix_elements_time',
              'non_sampled_scatter_time', 'eval_non_sampled_amps_time')

    def __init__(self, **kwargs):
        for f in self.FIELDS:
            setattr(self, f, kwargs.get(f, np.nan))

    def to_flat_dict(self):
        return {f: getattr(self, f) for f in self.FIELDS}

    def accumulate(self, other):
        for f in self.FIELDS:
            v, o = getattr(self, f), getattr(other, f)
            setattr(self, f, o if (isinstance(v, float) and np.isnan(v)) else v + o)


def parse_of_qubit_operator_arrays(of_qubit_operator, qubit_num: int):
    """PO:150-183 for single-word indices: (weights c128[T], xy int64[T], yz int64[T]) in dict order.
    Qubit q -> bit qubit_num-1-q (PO:162); bit 63 is the int64 sign bit (PO:164-167); each Y multiplies
    the weight by i (PO:176-177)."""
    arrays = getattr(of_qubit_operator, 'pauli_arrays', None)
    if arrays is not None:  # (xy, yz, w) already in X^x Z^z form: skips the python loop for 1e5+ terms
        xy, yz, w = arrays
        return (np.ascontiguousarray(w, dtype=np.complex128), np.ascontiguousarray(xy).view(np.int64).copy(),
                np.ascontiguousarray(yz).view(np.int64).copy())
    # one walk over the dict to flatten it, everything else vectorised (the reference loops in python over every Pauli of
    # every term with tensor indexing, PO:150-183)
    import itertools
    from operator import itemgetter
    terms = of_qubit_operator.terms
    T = len(terms)
    keys = list(terms.keys())
    w = np.fromiter(terms.values(), dtype=np.complex128, count=T)
    lens = np.fromiter(map(len, keys), dtype=np.int64, count=T)
    flat = list(itertools.chain.from_iterable(keys))
    m = len(flat)
    xy = np.zeros(T, np.uint64)
    yz = np.zeros(T, np.uint64)
    ny = np.zeros(T, np.int64)
    if m:
        q = np.fromiter(map(itemgetter(0), flat), dtype=np.int64, count=m)
        p = np.fromiter(map(ord, map(itemgetter(1), flat)), dtype=np.int64, count=m)
        bit = np.left_shift(np.uint64(1), (qubit_num - 1 - q).astype(np.uint64))
        is_y = p == ord('Y')
        is_x, is_z = (p == ord('X')) | is_y, (p == ord('Z')) | is_y
        ne = lens > 0
        starts = (np.cumsum(lens) - lens)[ne]
        xy[ne] = np.bitwise_or.reduceat(np.where(is_x, bit, np.uint64(0)), starts)
        yz[ne] = np.bitwise_or.reduceat(np.where(is_z, bit, np.uint64(0)), starts)
        ny[ne] = np.add.reduceat(is_y.astype(np.int64), starts)
    w = w * np.array([1.0 + 0j, 1j, -1.0 + 0j, -1j])[ny & 3]
    return w, xy.view(np.int64), yz.view(np.int64)


def count_qubits(of_qubit_operator) -> int:
    n = getattr(of_qubit_operator, 'qubit_num', None)
    if n is not None:
        return int(n)
    return 1 + max((q for t in of_qubit_operator.terms for q, _ in t), default=-1)


class PauliArraysOperator:
    """Minimal QubitOperator-like carrier for operators that already exist as (xy, yz, weight) arrays
    (what `synthetic.synthetic_hamiltonian` returns); `.terms` is built lazily for code that wants the dict."""

    def __init__(self, xy, yz, w, qubit_num):
        self.pauli_arrays = (np.asarray(xy), np.asarray(yz), np.asarray(w))
        self.qubit_num = qubit_num
        self._terms = None

    @property
    def terms(self):
        if self._terms is None:
            from .synthetic import pauli_arrays_to_terms
            xy, yz, w = self.pauli_arrays
            self._terms = pauli_arrays_to_terms(xy.view(np.uint64), yz.view(np.uint64), w, self.qubit_num)
        return self._terms

    def __len__(self):
        return int(self.pauli_arrays[0].shape[0])


class PauliObservable(AbstractHilbertSpaceObject):
    ALLOWED_COUPLING_METHODS = ('ham', 'all_to_all', 'hamming_ball', 'trie')
    MEMORY_MAGIC_CONSTANT = 25
    PAIR_JOIN_MIN_MASKS = 8192        # 'trie' / 'all_to_all' take the pair-join kernel for at least this many masks and N <= U / 2

    def __init__(self, *args, of_qubit_operator=None, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.qubit_num == count_qubits(of_qubit_operator)
        self.of_qubit_operator = of_qubit_operator
        self.term_num = len(of_qubit_operator) if isinstance(of_qubit_operator, PauliArraysOperator) else len(of_qubit_operator.terms)

        self.local_energy_structure_tensor_names = ('unq_xy_masks', 'unq_xy_masks_inv', 'unq_xy_to_yz_num',
                                                    'unq_xy_to_yz_start', 'rearranged_yz', 'rearranged_weights')
        self.tensor_name2tensor_path = {name: os.path.join(self.hilbert_space.parent_dir, f'{name}.npy')
                                        for name in self.local_energy_structure_tensor_names}
        if all(os.path.exists(p) for p in self.tensor_name2tensor_path.values()):  # PO:119-129
            host = {name: np.load(path) for name, path in self.tensor_name2tensor_path.items()}
        else:
            weights, xy_masks, yz_masks = parse_of_qubit_operator_arrays(of_qubit_operator, self.qubit_num)
            if self.device.type == 'cuda':  # sort / segment on the GPU (k2_sort.cu)
                host = self.compute_local_energy_structures_device(xy_masks, yz_masks, weights, self.device, self.hilbert_space._key_bits)
            else:  # host-side logic without a GPU (table tests, cache writers)
                host = self.compute_local_energy_structures_host(xy_masks, yz_masks, weights)
            os.makedirs(self.hilbert_space.parent_dir, exist_ok=True)
            for name, path in self.tensor_name2tensor_path.items():
                np.save(path, host[name])
        self._host_tables = host
        for name in self.local_energy_structure_tensor_names:
            setattr(self, name, pt.from_numpy(host[name]).to(self.device))
        self.unq_xy_masks_num = self.unq_xy_masks.shape[0]
        self._tables = None

    @staticmethod
    def compute_local_energy_structures_host(xy_masks, yz_masks, weights):
        """PO:131-142 + PO:185-211 without the python loops: unique XY masks in signed ascending order
        (torch.unique semantics), CSR (num, exclusive-cumsum start), YZ masks and weights regrouped with the
        original term order preserved inside each group."""
        unq, inv = np.unique(xy_masks, return_inverse=True)
        inv = inv.reshape(-1).astype(np.int64)
        num = np.bincount(inv, minlength=unq.shape[0]).astype(np.int64)
        start = np.cumsum(num) - num
        order = np.argsort(inv, kind='stable')
        return dict(unq_xy_masks=unq.reshape(-1, 1).astype(np.int64), unq_xy_masks_inv=inv,
                    unq_xy_to_yz_num=num, unq_xy_to_yz_start=start.astype(np.int64),
                    rearranged_yz=yz_masks[order].reshape(-1, 1).astype(np.int64),
                    rearranged_weights=np.ascontiguousarray(weights[order], dtype=np.complex128))

    @staticmethod
    def compute_local_energy_structures_device(xy_masks, yz_masks, weights, device, key_bits: int = 64):
        """The same six tensors built on the GPU (reference PO:131-142 + PO:185-211, two python
        loops with `.item()` per term): radix sort of the XY masks with head flags + scan for the unique masks and the inverse
        map (anqs_unique_i64), a stable radix sort of the inverse map for the regrouping permutation (terms of a group keep
        their original order), segment lengths from the sorted inverse map.  Returned as host arrays (they are cached as
        `.npy` files and handed to anqs_tables_create)."""
        dev = _lib.require_cuda(device)
        xy = pt.from_numpy(np.ascontiguousarray(xy_masks)).to(dev)
        yz = pt.from_numpy(np.ascontiguousarray(yz_masks)).to(dev)
        w = pt.from_numpy(np.ascontiguousarray(weights, dtype=np.complex128)).to(dev)
        unq, inv = _lib.unique_i64(xy, end_bit=key_bits)
        U = unq.shape[0]
        _, order = _lib.sort_pairs(inv, None, 0, max(1, int(U - 1).bit_length()))   # stable: original term order inside a group
        num = pt.bincount(inv, minlength=U)
        start = pt.cumsum(num, 0) - num
        return dict(unq_xy_masks=unq.reshape(-1, 1).cpu().numpy(), unq_xy_masks_inv=inv.cpu().numpy(),
                    unq_xy_to_yz_num=num.cpu().numpy(), unq_xy_to_yz_start=start.cpu().numpy(),
                    rearranged_yz=yz[order].reshape(-1, 1).cpu().numpy(), rearranged_weights=w[order].cpu().numpy())

    # ---- device handle ----------------------------------------------------------------------------------
    @property
    def tables(self):
        if self._tables is None:
            _lib.require_cuda(self.device)
            h = self._host_tables
            handle = _lib._vp()
            arrs = [np.ascontiguousarray(h['unq_xy_masks'].reshape(-1)), np.ascontiguousarray(h['unq_xy_to_yz_num']),
                    np.ascontiguousarray(h['unq_xy_to_yz_start']), np.ascontiguousarray(h['rearranged_yz'].reshape(-1)),
                    np.ascontiguousarray(h['rearranged_weights']).view(np.float64)]
            with pt.cuda.device(self.device):
                _lib.check(_lib.lib().anqs_tables_create(_lib.ctypes.byref(handle), self.qubit_num, arrs[0].shape[0],
                                                         arrs[3].shape[0], *[a.ctypes.data for a in arrs]))
                info = [_lib.ctypes.c_int(), _lib.ctypes.c_int64(), _lib.ctypes.c_int64(), _lib.ctypes.c_int(), _lib.ctypes.c_int64()]
                _lib.check(_lib.lib().anqs_tables_info(handle, *[_lib.ctypes.byref(i) for i in info]))
            self._tables = handle
            self._weights_real = bool(info[3].value)
            self._bitmap_row_words = int(info[4].value)
            self._enum_tiles = int(_lib.lib().anqs_k1_enum_tiles(handle))
        return self._tables

    # properties of the device tables: asking for one creates the tables (they used to be plain attributes that only
    # existed after the first kernel call)
    @property
    def mask_table(self) -> SampleTable:
        """Hash table {XY mask -> its index u} over the unique masks, for the pair-join kernel (built once)."""
        if getattr(self, '_mask_table', None) is None:
            self._mask_table = SampleTable(self.unq_xy_masks.to(self.device).contiguous().view(-1), None)
        return self._mask_table

    @property
    def weights_real(self) -> bool:
        """True when every Pauli weight has zero imaginary part (real-integral molecules): 8-byte matrix elements."""
        self.tables
        return self._weights_real

    @property
    def bitmap_row_words(self) -> int:
        self.tables
        return self._bitmap_row_words

    @property
    def enum_tiles(self) -> int:
        """Number of shared-memory tiles of the tiled enumeration (k1_enum.cu); 0 = the table does not fit it."""
        self.tables
        return self._enum_tiles

    def __del__(self):
        try:
            if getattr(self, '_tables', None) is not None:
                _lib.lib().anqs_tables_destroy(self._tables)
                self._tables = None
        except Exception:
            pass

    # ---- kernel 1: connected configurations -------------------------------------------------------------
    def connected_configurations(self, samples: pt.Tensor, alpha_num: int, beta_num: int, with_dest: bool = True,
                                 with_xy_ptr: bool = True, matrix_elements: str = None, tiled: bool = None, filter_variant: int = 0):
        """Connected list of a batch of packed samples [n] int64, lexicographic in (dest, xy_ptr) like
        PO:527-567.  matrix_elements in (None, 'real', 'complex').  Returns a dict with offsets [n+1] int64,
        dest int32 [M], xprime int64 [M], xy_ptr int32 [M], H.
        tiled: None = the tile-resident kernels (k1_enum.cu) whenever the table fits them, else the untiled pair
        (k1_connected.cu); True / False force one or the other.  Both give the same list bit for bit.
        filter_variant (tiled path): 0 = bit-sliced filter when the table allows it, 1 = product-layout filter."""
        tables = self.tables
        dev = self.device
        samples = samples.contiguous().view(-1)
        n = samples.shape[0]
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        if tiled is None:
            tiled = self.enum_tiles > 0
        elif tiled and self.enum_tiles == 0:
            raise RuntimeError('the tiled enumeration is unavailable for this Hamiltonian table')
        counts = pt.empty(n, dtype=pt.int64, device=dev)
        bitmap = pt.empty(n * self.bitmap_row_words, dtype=pt.int32, device=dev)
        enum_work = None
        if tiled:
            enum_work = pt.empty(max(1, (int(lib.anqs_k1_enum_workspace(tables, n)) + 3) // 4), dtype=pt.int32, device=dev)
            _lib.check(lib.anqs_k1_enum_filter_variant(tables, _lib.dptr(samples), n, alpha_num, beta_num, _lib.dptr(counts), _lib.dptr(bitmap),
                                                       _lib.dptr(enum_work), int(filter_variant), sp))
        else:
            _lib.check(lib.anqs_k1_filter(tables, _lib.dptr(samples), n, alpha_num, beta_num, _lib.dptr(counts), _lib.dptr(bitmap), sp))
        offsets = pt.empty(n + 1, dtype=pt.int64, device=dev)
        work = pt.empty(max(1, int(lib.anqs_scan_workspace(n)) // 8), dtype=pt.int64, device=dev)
        _lib.check(lib.anqs_exclusive_scan_i64(_lib.dptr(counts), _lib.dptr(offsets), n, _lib.dptr(work), sp))
        m = int(offsets[-1].item())  # the list is materialised, so its length has to come back to the host
        out = dict(offsets=offsets, counts=counts,
                   dest=pt.empty(m, dtype=pt.int32, device=dev) if with_dest else None,
                   xprime=pt.empty(m, dtype=pt.int64, device=dev),
                   xy_ptr=pt.empty(m, dtype=pt.int32, device=dev) if with_xy_ptr else None, H=None)
        hc = 0
        if matrix_elements == 'real':
            assert self.weights_real, 'Hamiltonian weights are complex'
            out['H'], hc = pt.empty(m, dtype=pt.float64, device=dev), 1
        elif matrix_elements == 'complex':
            out['H'], hc = pt.empty(m, dtype=pt.complex128, device=dev), 2
        h_ptr = _lib.dptr(pt.view_as_real(out['H'])) if hc == 2 else _lib.dptr(out['H'])
        if m == 0:
            return out
        if tiled:
            _lib.check(lib.anqs_k1_enum_emit(tables, _lib.dptr(samples), n, alpha_num, beta_num, _lib.dptr(bitmap), _lib.dptr(offsets), _lib.dptr(enum_work),
                                             _lib.dptr(out['dest']), _lib.dptr(out['xprime']), _lib.dptr(out['xy_ptr']), h_ptr, hc, sp))
        else:
            _lib.check(lib.anqs_k1_emit(tables, _lib.dptr(samples), n, _lib.dptr(bitmap), _lib.dptr(offsets), _lib.dptr(out['dest']),
                                        _lib.dptr(out['xprime']), _lib.dptr(out['xy_ptr']), h_ptr, hc, sp))
        return out

    # ---- reference method surface -----------------------------------------------------------------------
    def compute_matrix_elements(self, x_primes: pt.Tensor = None, ham_xy_pointers: pt.Tensor = None,
                                chunk_size: int = np.inf):
        """PO:255-324 (including the @timed tuple extension): (H [M] c128, yz_count, seconds)."""
        start = time.time()
        assert len(x_primes.shape) == 2
        assert len(ham_xy_pointers.shape) == 1
        assert x_primes.shape[0] == ham_xy_pointers.shape[0]
        m = x_primes.shape[0]
        H = pt.empty(m, dtype=pt.complex128, device=self.device)
        xp = x_primes.contiguous().view(-1)
        ptrs = ham_xy_pointers.to(pt.int64).contiguous()
        _lib.check(_lib.lib().anqs_matrix_elements(self.tables, _lib.dptr(xp), _lib.dptr(ptrs), m,
                                                   _lib.dptr(pt.view_as_real(H)), _lib.stream_ptr(self.device)))
        yz_num = int(self.unq_xy_to_yz_num[ptrs].sum().item()) if m > 0 else 0
        return H, yz_num, time.time() - start

    def find_sampled_and_coupled_via_ham(self, chunk_as_unq_batch_ptrs: pt.Tensor = None,
                                         unq_batch_as_base_indices: pt.Tensor = None, alpha_num: int = None,
                                         beta_num: int = None, metrics: LocalEnergyMetrics = None):
        """PO:569-600: (dest_as_chunk_ptrs, src_as_unq_batch_ptrs, src_as_base_indices [M,1],
        coupling_xy_as_unq_ham_xy_ptrs, metrics, seconds)."""
        start = time.time()
        metrics = metrics if metrics is not None else LocalEnergyMetrics()
        batch = unq_batch_as_base_indices.contiguous().view(-1)
        chunk = batch[chunk_as_unq_batch_ptrs]
        conn = self.connected_configurations(chunk, alpha_num, beta_num)
        metrics.candidate_x_primes_num = conn['xprime'].shape[0]
        mask, ptr = self.find_a_in_b(a=conn['xprime'].view(-1, 1), b=batch.view(-1, 1))
        metrics.candidates_time = metrics.filter_candidates_time = metrics.find_a_in_b_time = 0.0
        return (conn['dest'][mask].to(pt.int64), ptr[mask], conn['xprime'][mask].view(-1, 1),
                conn['xy_ptr'][mask].to(pt.int64), metrics, time.time() - start)

    def find_sampled_and_coupled(self, chunk_as_unq_batch_ptrs=None, unq_batch_as_base_indices=None,
                                 coupling_method: str = None, symmetric: bool = None, alpha_num: int = None,
                                 beta_num: int = None, metrics: LocalEnergyMetrics = None):
        assert coupling_method in self.ALLOWED_COUPLING_METHODS
        if coupling_method == 'hamming_ball':
            raise NotImplementedError("coupling_method='hamming_ball' is broken in the reference (pauli_observable.py:698)")
        return self.find_sampled_and_coupled_via_ham(chunk_as_unq_batch_ptrs=chunk_as_unq_batch_ptrs,
                                                     unq_batch_as_base_indices=unq_batch_as_base_indices,
                                                     alpha_num=alpha_num, beta_num=beta_num, metrics=metrics)

    @pt.no_grad()
    def compute_var_local_energy_proxy(self, unq_batch_as_base_indices: pt.Tensor = None,
                                       unq_batch_as_amps: pt.Tensor = None, coupling_method: str = None,
                                       chunk_size: int = 20000, alpha_num: int = None, beta_num: int = None,
                                       matrix_element_chunk_size: int = np.inf,
                                       row_start: int = 0, row_len: int = None,
                                       table: SampleTable = None, kernel_variant: int = 0,
                                       row_order: str = 'auto') -> Tuple[pt.Tensor, pt.Tensor, LocalEnergyMetrics]:
        """PO:396-487.  Sample-aware local energies of the whole batch in one fused launch; `chunk_size` and
        `matrix_element_chunk_size` are accepted for signature compatibility (nothing is materialised, so
        there is nothing to chunk).  row_start/row_len/table are extensions used by the multi-GPU shard path;
        kernel_variant: 0 = chosen from coupling_method and batch size, 1 = warp-per-sample kernel, 2 = bit-sliced kernel,
        3 = pair-join kernel (k1_pairs.cu).
        row_order: the fused kernels evaluate the rows in groups of 32 neighbours, and neighbours that resemble each other - rows in
        the sampler's tree order, or sorted - cost more than unrelated ones (measured at the C5 shape, 9.7e5 sampled rows: 22.6 ms
        in tree order, 21.5 sorted, 19.9 in a scattered order; profiles/README.md).  'strided' walks the rows in the order
        i -> i * P mod row_len (P ~ 0.618 row_len, coprime to it) and puts the results back where they belong; 'given' keeps the
        caller's order; 'auto' = 'strided' from 65 536 rows on."""
        assert coupling_method in self.ALLOWED_COUPLING_METHODS
        if coupling_method == 'hamming_ball':
            raise NotImplementedError("coupling_method='hamming_ball' is broken in the reference (pauli_observable.py:698)")
        dev = _lib.require_cuda(self.device)
        samples = unq_batch_as_base_indices.contiguous().view(-1)
        amps = unq_batch_as_amps.to(pt.complex128).contiguous()
        n = samples.shape[0]
        assert amps.shape[0] == n
        row_len = n - row_start if row_len is None else row_len
        eloc = pt.empty(row_len, dtype=pt.complex128, device=dev)
        # 'trie' / 'all_to_all' couple the sampled configurations with each other (PO:602-696): the pair-join kernel, N^2 pair
        # tests instead of N x U filter tests, while that is the smaller number; 'ham' (and large batches) enumerate and probe
        # (measured on B200, profiles/README.md: 1.8x / 1.55x faster than enumerate-and-probe at 1e3 / 1e4 samples against the
        # 23 157 masks of C5, slower against the 2 536 masks of the dense 20-qubit shape and beyond ~U/2 samples)
        pair_join = kernel_variant == 3 or (kernel_variant == 0 and coupling_method in ('trie', 'all_to_all')
                                            and self.unq_xy_masks_num >= self.PAIR_JOIN_MIN_MASKS and 2 * n <= self.unq_xy_masks_num)
        if pair_join:
            mt = self.mask_table
            work = _lib._workspace(_lib.lib().anqs_pair_join_workspace(row_len, n), dev)
            _lib.check(_lib.lib().anqs_local_energy_pair_join(
                self.tables, _lib.dptr(samples), _lib.dptr(pt.view_as_real(amps)), n, row_start, row_len, _lib.dptr(mt.slots), mt.capacity,
                alpha_num, beta_num, _lib.dptr(pt.view_as_real(eloc)), _lib.dptr(work), _lib.stream_ptr(dev)))
            return eloc, eloc, LocalEnergyMetrics()
        if table is None:
            table = SampleTable(samples, amps)
        assert row_order in ('auto', 'given', 'strided')
        if row_order == 'strided' or (row_order == 'auto' and row_len >= 65536):
            # the table holds the whole sampled set; the rows of this call are passed as their own (permuted) array
            step = int(row_len * 0.6180339887) | 1
            while math.gcd(step, row_len) != 1:
                step += 2
            order = (pt.arange(row_len, dtype=pt.int64, device=dev) * step) % row_len
            rows = samples[row_start:row_start + row_len][order]
            row_amps = amps[row_start:row_start + row_len][order]
            e_rows = pt.empty(row_len, dtype=pt.complex128, device=dev)
            _lib.check(_lib.lib().anqs_local_energy_sample_aware_variant(
                self.tables, _lib.dptr(rows), _lib.dptr(pt.view_as_real(row_amps)), row_len, 0, row_len,
                _lib.dptr(table.slots), table.capacity, alpha_num, beta_num, _lib.dptr(pt.view_as_real(e_rows)), int(kernel_variant),
                _lib.stream_ptr(dev)))
            eloc[order] = e_rows
            return eloc, eloc, LocalEnergyMetrics()
        _lib.check(_lib.lib().anqs_local_energy_sample_aware_variant(
            self.tables, _lib.dptr(samples), _lib.dptr(pt.view_as_real(amps)), n, row_start, row_len,
            _lib.dptr(table.slots), table.capacity, alpha_num, beta_num, _lib.dptr(pt.view_as_real(eloc)), int(kernel_variant),
            _lib.stream_ptr(dev)))
        return eloc, eloc, LocalEnergyMetrics()

    @pt.no_grad()
    def compute_local_energies(self, wf=None, sampled_indices: pt.Tensor = None, sampled_amps: pt.Tensor = None,
                               verbose: bool = False, use_tree_for_candidates: bool = False, chunk_size: int = 20000,
                               sample_aware: bool = False, compute_via_ham_xy_coupling: bool = True,
                               amps_chunk_size: int = 100000) -> Tuple[pt.Tensor, pt.Tensor, LocalEnergyMetrics]:
        """PO:326-393 -> PO:992-1105.  Returns (full E_loc, sample-aware E_loc, metrics).  With
        sample_aware=True both are the sample-aware value (as in the reference); otherwise connected
        configurations outside the sampled set are de-duplicated, evaluated with wf.amplitude and added."""
        alpha_num = beta_num = wf.masker.symmetries[0].particle_num // 2  # PO:979,985 (closed shell)
        samples = sampled_indices.contiguous().view(-1)
        amps = sampled_amps.to(pt.complex128).contiguous()
        n = samples.shape[0]
        table = SampleTable(samples, amps)
        if sample_aware:
            return self.compute_var_local_energy_proxy(unq_batch_as_base_indices=sampled_indices, unq_batch_as_amps=amps,
                                                       coupling_method='ham', alpha_num=alpha_num, beta_num=beta_num, table=table)
        dev = self.device
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        metrics = LocalEnergyMetrics()
        full = pt.empty(n, dtype=pt.complex128, device=dev)
        aware = pt.empty(n, dtype=pt.complex128, device=dev)
        amps_r = pt.view_as_real(amps)
        for lo in range(0, n, chunk_size):
            hi = min(n, lo + chunk_size)
            cur = LocalEnergyMetrics()
            conn = self.connected_configurations(samples[lo:hi], alpha_num, beta_num, with_dest=False, with_xy_ptr=False,
                                                 matrix_elements='real' if self.weights_real else 'complex')
            xp, H, offsets = conn['xprime'], conn['H'], conn['offsets']
            hc = 1 if self.weights_real else 2
            h_ptr = _lib.dptr(H) if hc == 1 else _lib.dptr(pt.view_as_real(H))
            m = xp.shape[0]
            cur.candidate_x_primes_num = (hi - lo) * self.unq_xy_masks_num  # pre-filter count, as PO:1006 reports it
            ptr = pt.empty(m, dtype=pt.int64, device=dev)
            _lib.check(lib.anqs_hash_probe(_lib.dptr(table.slots), table.capacity, _lib.dptr(xp), m, _lib.dptr(ptr), _lib.dptr(None), sp))
            # sampled part: E_s[i] = sum H psi(x') / psi(x_i)   (PO:1048-1057)
            aware_chunk = pt.empty(hi - lo, dtype=pt.complex128, device=dev)
            dest_amps = amps_r[lo:hi].contiguous()
            _lib.check(lib.anqs_accumulate_rows(_lib.dptr(offsets), hi - lo, _lib.dptr(ptr), h_ptr, hc, _lib.dptr(amps_r),
                                                _lib.dptr(dest_amps), _lib.dptr(pt.view_as_real(aware_chunk)), 0, sp))
            aware[lo:hi] = aware_chunk
            # non-sampled part (PO:1062-1103): unique -> amplitudes -> gather back through the inverse map
            missing = ptr < 0
            cur.sampled_x_primes_num = int((~missing).sum().item())
            cur.non_sampled_x_primes_num = m - cur.sampled_x_primes_num
            if cur.non_sampled_x_primes_num > 0:
                unq, inv = _lib.unique_i64(xp[missing], end_bit=self.hilbert_space._key_bits)  # PO:1016-1040 (k2_sort.cu)
                cur.non_sampled_unq_x_primes_num = unq.shape[0]
                t0 = time.time()
                unq_amps = pt.cat([wf.amplitude(unq[a:a + amps_chunk_size].view(-1, 1)).detach()
                                   for a in range(0, unq.shape[0], amps_chunk_size)]).to(pt.complex128).contiguous()
                cur.eval_non_sampled_amps_time = time.time() - t0
                inv_full = pt.full((m,), -1, dtype=pt.int64, device=dev)
                inv_full[missing] = inv
                rest = pt.empty(hi - lo, dtype=pt.complex128, device=dev)
                _lib.check(lib.anqs_accumulate_rows(_lib.dptr(offsets), hi - lo, _lib.dptr(inv_full), h_ptr, hc,
                                                    _lib.dptr(pt.view_as_real(unq_amps)), _lib.dptr(dest_amps),
                                                    _lib.dptr(pt.view_as_real(rest)), 0, sp))
                full[lo:hi] = aware_chunk + rest
            else:
                # the reference raises IndexError here (expand_pointers on an empty list);
                # the mathematically correct answer is "nothing to add"
                cur.non_sampled_unq_x_primes_num = 0
                full[lo:hi] = aware_chunk
            metrics.accumulate(cur)
        return full, aware, metrics
