"""TEST INFRASTRUCTURE ONLY — numpy/ctypes front end of the C oracle (oracle/anqs_oracle.c) plus an
independent dense-matrix checker.  The product never imports this module.

Parity status: pinned against the unmodified reference (tests/golden/*.npz, made by
oracle/make_golden.py through oracle/ref_shim.py); see the header of anqs_oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, 'libanqs_oracle.so')
    src = os.path.join(_HERE, 'anqs_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'libanqs_oracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_build_tables.restype = ctypes.c_int64
        _LIB.orc_candidates_ham.restype = ctypes.c_int64
        _LIB.orc_unique.restype = ctypes.c_int64
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _i64(a):
    return np.ascontiguousarray(np.asarray(a).view(np.int64) if np.asarray(a).dtype == np.uint64 else a, dtype=np.int64)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def popcount(a):
    a = _i64(a).reshape(-1)
    out = np.empty_like(a)
    lib().orc_popcount(_p(a, _i64p), _p(out, _i64p), ctypes.c_int64(a.size))
    return out


class Tables:
    """The six local-energy structure tensors of pauli_observable.py:110-115."""

    def __init__(self, xy, yz, w):
        xy, yz = _i64(xy).reshape(-1), _i64(yz).reshape(-1)
        w = np.ascontiguousarray(w, dtype=np.complex128).reshape(-1)
        T = xy.size
        unq = np.empty(T, np.int64); inv = np.empty(T, np.int64)
        num = np.empty(T, np.int64); start = np.empty(T, np.int64)
        re_yz = np.empty(T, np.int64); re_w = np.empty(T, np.complex128)
        U = lib().orc_build_tables(_p(xy, _i64p), _p(yz, _i64p), _p(w.view(np.float64), _f64p), ctypes.c_int64(T),
                                   _p(unq, _i64p), _p(inv, _i64p), _p(num, _i64p), _p(start, _i64p),
                                   _p(re_yz, _i64p), _p(re_w.view(np.float64), _f64p))
        self.term_num = T
        self.unq_xy_masks_num = int(U)
        self.unq_xy_masks = unq[:U].copy()
        self.unq_xy_masks_inv = inv
        self.unq_xy_to_yz_num = num[:U].copy()
        self.unq_xy_to_yz_start = start[:U].copy()
        self.rearranged_yz = re_yz
        self.rearranged_weights = re_w


def terms_to_arrays(terms: dict, qubit_num: int):
    """pauli_observable.py:150-183 parse_of_qubit_operator for one-word indices: qubit q -> bit n-1-q,
    bit 63 carried as the int64 sign bit, every Y multiplies the weight by i."""
    T = len(terms)
    xy = np.zeros(T, np.uint64); yz = np.zeros(T, np.uint64); w = np.zeros(T, np.complex128)
    for t, (ops, c) in enumerate(terms.items()):
        x = z = 0
        c = c + 0j
        for q, p in ops:
            bit = 1 << (qubit_num - 1 - q)
            if p in ('X', 'Y'):
                x |= bit
            if p in ('Y', 'Z'):
                z |= bit
            if p == 'Y':
                c *= 1j
        xy[t], yz[t], w[t] = x, z, c
    return xy.view(np.int64), yz.view(np.int64), w


def candidates_ham(samples, chunk_start, chunk_len, tab: Tables, alpha_num, beta_num):
    s = _i64(samples).reshape(-1)
    args = (_p(s, _i64p), ctypes.c_int64(chunk_start), ctypes.c_int64(chunk_len), _p(tab.unq_xy_masks, _i64p),
            ctypes.c_int64(tab.unq_xy_masks_num), ctypes.c_int64(alpha_num), ctypes.c_int64(beta_num))
    m = lib().orc_candidates_ham(*args, None, None, None)
    dest = np.empty(m, np.int64); xp = np.empty(m, np.int64); ptr = np.empty(m, np.int64)
    lib().orc_candidates_ham(*args, _p(dest, _i64p), _p(xp, _i64p), _p(ptr, _i64p))
    return dest, xp, ptr


def find_a_in_b(a, b):
    a, b = _i64(a).reshape(-1), _i64(b).reshape(-1)
    mask = np.empty(a.size, np.uint8); ptr = np.empty(a.size, np.int64)
    lib().orc_find_a_in_b(_p(a, _i64p), ctypes.c_int64(a.size), _p(b, _i64p), ctypes.c_int64(b.size),
                          _p(mask, _u8p), _p(ptr, _i64p))
    return mask.astype(bool), ptr


def matrix_elements(xprime, xy_ptr, tab: Tables):
    xp, ptr = _i64(xprime).reshape(-1), _i64(xy_ptr).reshape(-1)
    H = np.empty(xp.size, np.complex128)
    lib().orc_matrix_elements(_p(xp, _i64p), _p(ptr, _i64p), ctypes.c_int64(xp.size),
                              _p(tab.unq_xy_to_yz_start, _i64p), _p(tab.unq_xy_to_yz_num, _i64p),
                              _p(tab.rearranged_yz, _i64p), _p(tab.rearranged_weights.view(np.float64), _f64p),
                              _p(H.view(np.float64), _f64p))
    return H


class SampledSet:
    """The {configuration -> position} map of a sampled set, built once and reused by every row window evaluated
    against it (oracle/anqs_oracle.c orc_map_create)."""

    def __init__(self, samples):
        self.samples = _i64(samples).reshape(-1)
        lib().orc_map_create.restype = ctypes.c_void_p
        self.handle = ctypes.c_void_p(lib().orc_map_create(_p(self.samples, _i64p), ctypes.c_int64(self.samples.size)))

    def __del__(self):
        try:
            if self.handle:
                lib().orc_map_free(self.handle)
                self.handle = None
        except Exception:
            pass


def local_energy_sample_aware(samples, amps, tab: Tables, alpha_num, beta_num, row_start=0, row_len=None, sampled_set: SampledSet = None):
    s = _i64(samples).reshape(-1) if sampled_set is None else sampled_set.samples
    a = np.ascontiguousarray(amps, dtype=np.complex128).reshape(-1)
    row_len = s.size - row_start if row_len is None else row_len
    e = np.empty(row_len, np.complex128)
    head = (_p(s, _i64p), _p(a.view(np.float64), _f64p), ctypes.c_int64(s.size))
    tail = (ctypes.c_int64(row_start),
            ctypes.c_int64(row_len), _p(tab.unq_xy_masks, _i64p), ctypes.c_int64(tab.unq_xy_masks_num),
            _p(tab.unq_xy_to_yz_start, _i64p), _p(tab.unq_xy_to_yz_num, _i64p), _p(tab.rearranged_yz, _i64p),
            _p(tab.rearranged_weights.view(np.float64), _f64p), ctypes.c_int64(alpha_num), ctypes.c_int64(beta_num),
            _p(e.view(np.float64), _f64p))
    if sampled_set is None:
        lib().orc_local_energy_sample_aware(*head, *tail)
    else:
        lib().orc_local_energy_sample_aware_map(*head, sampled_set.handle, *tail)
    return e


def sort_base_idx(a):
    a = _i64(a).reshape(-1)
    s = np.empty_like(a); p = np.empty_like(a)
    lib().orc_sort_base_idx(_p(a, _i64p), ctypes.c_int64(a.size), _p(s, _i64p), _p(p, _i64p))
    return s, p


def unique(a):
    a = _i64(a).reshape(-1)
    u = np.empty_like(a); inv = np.empty_like(a)
    n = lib().orc_unique(_p(a, _i64p), ctypes.c_int64(a.size), _p(u, _i64p), _p(inv, _i64p))
    return u[:n].copy(), inv


# ---------------------------------------------------------------------------------------------
# Independent dense-matrix checker (Kronecker products; qubit 0 = most significant bit, which is
# the packed-index convention of pauli_observable.py:162).  n <= ~14.
# ---------------------------------------------------------------------------------------------
_PAULI = {
    'I': np.eye(2, dtype=np.complex128),
    'X': np.array([[0, 1], [1, 0]], dtype=np.complex128),
    'Y': np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    'Z': np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def dense_matrix_from_terms(terms: dict, qubit_num: int) -> np.ndarray:
    dim = 1 << qubit_num
    H = np.zeros((dim, dim), dtype=np.complex128)
    for ops, c in terms.items():
        letters = ['I'] * qubit_num
        for q, p in ops:
            letters[q] = p
        m = np.array([[1.0 + 0j]])
        for l in letters:
            m = np.kron(m, _PAULI[l])
        H += c * m
    return H


def dense_matrix_from_arrays(xy, yz, w, qubit_num: int) -> np.ndarray:
    """Same matrix from the X^x Z^z representation (fast path for n <= 14)."""
    dim = 1 << qubit_num
    H = np.zeros((dim, dim), dtype=np.complex128)
    xs = np.arange(dim, dtype=np.uint64)
    for a, b, c in zip(np.asarray(xy).view(np.uint64), np.asarray(yz).view(np.uint64), w):
        sign = 1.0 - 2.0 * (np.bitwise_count(xs & b) & 1)
        H[(xs ^ a).astype(np.int64), xs.astype(np.int64)] += c * sign
    return H


def dense_local_energy(H: np.ndarray, samples, amps) -> np.ndarray:
    """Sample-aware E_loc[i] = (H[S,S] psi)[i] / psi[i]."""
    s = np.asarray(samples).view(np.uint64).astype(np.int64)
    return (H[np.ix_(s, s)] @ amps) / amps
