"""HilbertSpace drop-in (reference: nqs/nqs/base/hilbert_space.py:9-284).

Same constructor keywords and method names as the reference so that PauliObservable, the ANQS modules and
the reference's own `experiments/calculations/*` call sites work unchanged.  Occupation bitstrings are
packed one int64 per configuration; only single-word indices (qubit_num <= 64, HS:53 int_per_idx == 1) are
supported, which covers every BASELINE.json configuration.

Compute methods run on the GPU through libanqs_b200.so; there is no CPU path.  Pure host logic
(constructor, shapes, permutations) works without a GPU.
"""
import numpy as np
import torch as pt

from . import _lib
from .constants import BASE_INT_TYPE, BASE_REAL_TYPE, BASE_COMPLEX_TYPE


class HilbertSpace:
    SUPPORTED_IDX_DTYPES = (pt.int64,)
    SUPPORTED_RDTYPES = (pt.double,)
    SUPPORTED_CDTYPES = (pt.cdouble,)
    DEFAULT_GPU_MEMORY_LIMIT = 6 * 10 ** 9
    ALLOWED_PERM_TYPES = ('direct', 'inverse')
    # the reference's three modes select *how* popcount is computed (HS:106-117); here all of them map to
    # the same POPC-instruction kernel
    ALLOWED_POPCOUNT_MODES = ('compute_efficient', 'memory_efficient', 'custom')

    def __init__(self, *, qubit_num: int = 0, device=None, rdtype=BASE_REAL_TYPE, cdtype=BASE_COMPLEX_TYPE,
                 idx_dtype=BASE_INT_TYPE, parent_dir: str = None, rng_seed: int = None, rng=None,
                 gpu_memory_limit: int = DEFAULT_GPU_MEMORY_LIMIT, perm_type: str = 'direct',
                 popcount_mode: str = 'custom'):
        assert idx_dtype in HilbertSpace.SUPPORTED_IDX_DTYPES
        self.idx_dtype = idx_dtype
        assert device is not None
        self.device = pt.device(device)
        self.qubit_num = qubit_num
        self.bit_depth = 64
        self.int_per_idx = (self.qubit_num // self.bit_depth) + 1 * ((self.qubit_num % self.bit_depth) > 0)
        if self.int_per_idx != 1:
            raise NotImplementedError('anqs_b200 supports single-word indices only (1 <= qubit_num <= 64)')
        assert rdtype in HilbertSpace.SUPPORTED_RDTYPES
        self.rdtype = rdtype
        assert cdtype in HilbertSpace.SUPPORTED_CDTYPES
        self.cdtype = cdtype
        assert parent_dir is not None
        self.parent_dir = parent_dir
        assert rng_seed is not None
        self.rng_seed = rng_seed
        if rng is None:
            self.rng = np.random.default_rng(seed=self.rng_seed)
            pt.manual_seed(self.rng_seed)  # HS:90
        else:
            self.rng = rng
        self.gpu_memory_limit = gpu_memory_limit
        self.max_idx_num_per_mask = self.gpu_memory_limit // (8 * self.qubit_num)
        assert perm_type in self.ALLOWED_PERM_TYPES
        self.perm_type = perm_type
        self.init_perm(perm_type)
        assert popcount_mode in self.ALLOWED_POPCOUNT_MODES
        self.popcount_mode = popcount_mode
        self.shifts = pt.arange(0, self.qubit_num, dtype=self.idx_dtype, device=self.device)

    def init_perm(self, perm_type: str = 'direct'):
        assert perm_type in self.ALLOWED_PERM_TYPES
        if perm_type == 'direct':
            self.perm = pt.arange(self.qubit_num, dtype=self.idx_dtype, device=self.device)
            self.inv_perm = pt.arange(self.qubit_num, dtype=self.idx_dtype, device=self.device)
        else:
            self.perm = pt.arange(self.qubit_num - 1, -1, -1, dtype=self.idx_dtype, device=self.device)
            self.inv_perm = pt.arange(self.qubit_num - 1, -1, -1, dtype=self.idx_dtype, device=self.device)
        return self.perm, self.inv_perm

    # ---- codec (HS:121-147) ------------------------------------------------------------------------
    def base_idx2base_vec(self, base_idx: pt.Tensor) -> pt.Tensor:
        if not pt.is_tensor(base_idx):
            base_idx = pt.tensor(base_idx, dtype=self.idx_dtype, device=self.device)
        assert len(base_idx.shape) == 2
        assert base_idx.shape[-1] == self.int_per_idx
        assert base_idx.device == self.device
        return (base_idx[:, 0].reshape(-1, 1) >> self.shifts).bitwise_and_(1)

    def base_vec2base_idx(self, base_vec: pt.Tensor) -> pt.Tensor:
        if not pt.is_tensor(base_vec):
            base_vec = pt.tensor(base_vec, dtype=self.idx_dtype, device=self.device)
        if base_vec.dtype != self.idx_dtype:
            base_vec = base_vec.type(self.idx_dtype)
        assert base_vec.device == self.device
        return pt.sum(base_vec << self.shifts[:base_vec.shape[-1]], dim=-1, keepdim=True)

    # ---- popcount (HS:158-192 -> POPC:34-87) -------------------------------------------------------
    def popcount(self, base_idx: pt.Tensor) -> pt.Tensor:
        _lib.require_cuda(base_idx.device)
        assert base_idx.dtype == pt.int64
        src = base_idx.contiguous().view(-1)
        out = pt.empty_like(src)
        _lib.check(_lib.lib().anqs_popcount_i64(_lib.dptr(src), _lib.dptr(out), src.numel(), _lib.stream_ptr(src.device)))
        return out.view(-1, self.int_per_idx).sum(dim=-1) if self.int_per_idx > 1 else out

    def popcount_(self, base_idx: pt.Tensor) -> pt.Tensor:
        _lib.require_cuda(base_idx.device)
        assert base_idx.dtype == pt.int64
        if not base_idx.is_contiguous():
            return self.popcount(base_idx)
        flat = base_idx.view(-1)
        _lib.check(_lib.lib().anqs_popcount_i64(_lib.dptr(flat), _lib.dptr(flat), flat.numel(), _lib.stream_ptr(flat.device)))
        return flat

    def old_popcount(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.popcount(base_idx)

    # ---- unique / sort / join (HS:200-284) -----------------------------------------------------------
    @property
    def _key_bits(self) -> int:
        """Bits of a packed index that can differ between two configurations (all of them at 64 qubits)."""
        return 64 if self.qubit_num >= 64 else self.qubit_num

    def compute_unique_indices(self, base_idx):
        """Sorted (signed) unique rows and the inverse map (HS:215-228): radix sort + head flags + scan (k2_sort.cu)."""
        assert len(base_idx.shape) == 2
        assert base_idx.shape[-1] == self.int_per_idx
        unq, inv = _lib.unique_i64(base_idx[..., 0], end_bit=self._key_bits)
        return unq.reshape(-1, 1), inv

    def sort_base_idx(self, base_idx: pt.Tensor = None, descending: bool = False):
        """Ascending in UNSIGNED order with a stable permutation (HS:239-261): LSD radix sort over the qubit_num low bits."""
        if descending:
            raise NotImplementedError
        srt, perm = _lib.sort_pairs(base_idx[:, 0], None, 0, self._key_bits)
        return srt.view(-1, 1), perm

    def find_a_in_b(self, a: pt.Tensor, b: pt.Tensor):
        """(mask, ptr): position of each row of `a` in `b`, -1 when absent (HS:263-284)."""
        assert len(a.shape) <= 2 and len(b.shape) <= 2 and len(a.shape) == len(b.shape)
        if len(a.shape) == 2:
            assert a.shape[1] == b.shape[1] == self.int_per_idx
        dev = _lib.require_cuda(a.device)
        a_flat = a.contiguous().view(-1)
        b_flat = b.contiguous().view(-1)
        table = SampleTable(b_flat, None)
        ptr = pt.empty(a_flat.shape[0], dtype=pt.int64, device=dev)
        mask = pt.empty(a_flat.shape[0], dtype=pt.uint8, device=dev)
        _lib.check(_lib.lib().anqs_hash_probe(_lib.dptr(table.slots), table.capacity, _lib.dptr(a_flat), a_flat.shape[0],
                                              _lib.dptr(ptr), _lib.dptr(mask), _lib.stream_ptr(dev)))
        return mask.bool(), ptr


class SampleTable:
    """Open-addressing table {configuration -> (position, amplitude)} plus its presence filter, in device memory."""

    def __init__(self, keys: pt.Tensor, amps: pt.Tensor = None, spread_bits: int = None):
        dev = _lib.require_cuda(keys.device)
        n = keys.shape[0]
        self.capacity = int(_lib.lib().anqs_hash_capacity(n))
        nbytes = int(_lib.lib().anqs_hash_bytes(self.capacity))
        self.slots = pt.empty(((nbytes + 7) // 8,), dtype=pt.int64, device=dev)  # slots + header + filter
        assert self.slots.data_ptr() % 128 == 0
        self.rebuild(keys, amps, spread_bits)

    def rebuild(self, keys: pt.Tensor, amps: pt.Tensor = None, spread_bits: int = None):
        """(Re)builds the table in place for a new key set that needs the same capacity: an iteration loop keeps one table
        object and pays no allocation per batch (stream-ordered: four kernels and three memsets, no host synchronisation)."""
        dev = _lib.require_cuda(keys.device)
        assert keys.dtype == pt.int64 and keys.dim() == 1 and keys.is_contiguous() and keys.device == self.slots.device
        n = keys.shape[0]
        if int(_lib.lib().anqs_hash_capacity(n)) != self.capacity:
            raise ValueError(f'SampleTable.rebuild: {n} keys need capacity {int(_lib.lib().anqs_hash_capacity(n))}, this table has {self.capacity}')
        self.n = n
        amps_real = None
        if amps is not None:
            assert amps.dtype == pt.complex128 and amps.shape[0] == n
            amps_real = pt.view_as_real(amps.contiguous())
        self._keep = (keys, amps_real)
        if spread_bits is None:
            _lib.check(_lib.lib().anqs_hash_build(_lib.dptr(keys), _lib.dptr(amps_real), n, _lib.dptr(self.slots), self.capacity,
                                                  _lib.stream_ptr(dev)))
        else:
            _lib.check(_lib.lib().anqs_hash_build_spread(_lib.dptr(keys), _lib.dptr(amps_real), n, _lib.dptr(self.slots),
                                                         self.capacity, int(spread_bits), _lib.stream_ptr(dev)))
        return self

    def filter_info(self):
        """(G, overloaded[0..6]): the spread chosen by the build and the line-occupancy statistic behind it."""
        import ctypes
        g = ctypes.c_int(0)
        over = (ctypes.c_int64 * 7)()
        _lib.check(_lib.lib().anqs_hash_filter_info(_lib.dptr(self.slots), self.capacity, ctypes.byref(g), over,
                                                    _lib.stream_ptr(self.slots.device)))
        return int(g.value), [int(v) for v in over]
