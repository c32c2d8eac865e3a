"""ctypes binding of libanqs_b200.so (C ABI in include/anqs_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller gets a
RuntimeError.  PyTorch is used only to own device memory and streams; every signature below is plain
pointers and sizes.
"""
import ctypes
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, 'libanqs_b200.so')

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_vp = ctypes.c_void_p

_SIGNATURES = {
    'anqs_abi_version': (_c_int, []),
    'anqs_last_error': (ctypes.c_char_p, []),
    'anqs_device_check': (_c_int, [_c_int, _vp, _vp, _vp]),
    'anqs_popcount_i64': (_c_int, [_vp, _vp, _c_i64, _vp]),
    'anqs_tables_create': (_c_int, [_vp, _c_int, _c_i64, _c_i64, _vp, _vp, _vp, _vp, _vp]),
    'anqs_tables_destroy': (_c_int, [_vp]),
    'anqs_tables_info': (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_k1_filter': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp]),
    'anqs_scan_workspace': (ctypes.c_size_t, [_c_i64]),
    'anqs_exclusive_scan_i64': (_c_int, [_vp, _vp, _c_i64, _vp, _vp]),
    'anqs_k1_emit': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp]),
    'anqs_k1_enum_tiles': (_c_int, [_vp]),
    'anqs_k1_enum_workspace': (ctypes.c_size_t, [_vp, _c_i64]),
    'anqs_k1_enum_filter': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp]),
    'anqs_k1_enum_filter_variant': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp]),
    'anqs_k1_enum_emit': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp]),
    'anqs_matrix_elements': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    'anqs_hash_capacity': (_c_i64, [_c_i64]),
    'anqs_hash_bytes': (ctypes.c_size_t, [_c_i64]),
    'anqs_hash_build': (_c_int, [_vp, _vp, _c_i64, _vp, _c_i64, _vp]),
    'anqs_hash_build_spread': (_c_int, [_vp, _vp, _c_i64, _vp, _c_i64, _c_int, _vp]),
    'anqs_hash_filter_info': (_c_int, [_vp, _c_i64, _vp, _vp, _vp]),
    'anqs_hash_probe': (_c_int, [_vp, _c_i64, _vp, _c_i64, _vp, _vp, _vp]),
    'anqs_local_energy_sample_aware': (_c_int, [_vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp, _vp]),
    'anqs_local_energy_sample_aware_variant': (_c_int, [_vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp, _c_int, _vp]),
    'anqs_accumulate_rows': (_c_int, [_vp, _c_i64, _vp, _vp, _c_int, _vp, _vp, _vp, _c_int, _vp]),
    'anqs_made_log_psi': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    'anqs_made_cond_log_abs': (_c_int, [_vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_made_backward_chain': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_batch_reduce_workspace': (_c_i64, [_vp, _c_int, _c_i64]),
    'anqs_batch_reduce_gemm': (_c_int, [_vp, _c_int, _c_i64, _c_int, _vp, _c_i64, _vp]),
    'anqs_nade_log_psi': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    'anqs_nade_backward_chain': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_nade_cond_log_abs': (_c_int, [_vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_nade_tc_packed_bytes': (ctypes.c_size_t, [_vp]),
    'anqs_nade_tc_pack': (_c_int, [_vp, _vp, _vp]),
    'anqs_nade_log_psi_tc': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    'anqs_nade_cond_log_abs_tc': (_c_int, [_vp, _vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_made_tc_packed_bytes': (ctypes.c_size_t, [_vp]),
    'anqs_made_tc_pack': (_c_int, [_vp, _vp, _vp]),
    'anqs_made_log_psi_tc': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    'anqs_made_cond_log_abs_tc': (_c_int, [_vp, _vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_transformer_log_psi': (_c_int, [_vp, _vp, _c_i64, _vp, _vp]),
    'anqs_transformer_cond_log_abs': (_c_int, [_vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_transformer_tc_packed_bytes': (ctypes.c_size_t, [_vp]),
    'anqs_transformer_tc_pack': (_c_int, [_vp, _vp, _vp]),
    'anqs_transformer_log_psi_tc': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    'anqs_transformer_cond_log_abs_tc': (_c_int, [_vp, _vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_sampler_split_level': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _c_i64, _c_i64, _c_int, _c_int, ctypes.c_uint64,
                                          _c_i64, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_emit_children': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_gumbel_select': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_gumbel_level': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _c_int, ctypes.c_uint64,
                                           _c_i64, _vp, _vp, _vp, _vp]),
}


class MadeDesc(ctypes.Structure):
    """anqs_made_desc_t (include/anqs_b200.h)."""
    _fields_ = [('qubit_num', ctypes.c_int32), ('qudit_num', ctypes.c_int32), ('max_qudit_dim', ctypes.c_int32),
                ('depth', ctypes.c_int32), ('width', ctypes.c_int32), ('use_res', ctypes.c_int32),
                ('subtract_mean', ctypes.c_int32), ('sym_num', ctypes.c_int32),
                ('qudit_starts', ctypes.c_int32 * 65), ('du', ctypes.c_uint8 * 64), ('sym', (ctypes.c_int64 * 8) * 8),
                ('w_abs', ctypes.c_void_p * 5), ('b_abs', ctypes.c_void_p * 5), ('w_phase', ctypes.c_void_p * 5),
                ('b_phase', ctypes.c_void_p * 5), ('cont_mask', ctypes.c_void_p), ('memo_size', ctypes.c_int64)]

class NadeDesc(ctypes.Structure):
    """anqs_nade_desc_t (include/anqs_b200.h)."""
    _fields_ = [('qubit_num', ctypes.c_int32), ('qudit_num', ctypes.c_int32), ('max_qudit_dim', ctypes.c_int32),
                ('depth', ctypes.c_int32), ('width', ctypes.c_int32), ('use_res', ctypes.c_int32),
                ('subtract_mean', ctypes.c_int32), ('sym_num', ctypes.c_int32),
                ('qudit_starts', ctypes.c_int32 * 65), ('du', ctypes.c_uint8 * 64), ('sym', (ctypes.c_int64 * 8) * 8),
                ('ptrs', ctypes.c_void_p), ('cont_mask', ctypes.c_void_p), ('memo_size', ctypes.c_int64)]


class TransformerDesc(ctypes.Structure):
    """anqs_transformer_desc_t (include/anqs_b200.h)."""
    _fields_ = [('qubit_num', ctypes.c_int32), ('dim', ctypes.c_int32), ('depth', ctypes.c_int32), ('head_num', ctypes.c_int32),
                ('sym_num', ctypes.c_int32), ('pad0', ctypes.c_int32), ('pad1', ctypes.c_int32), ('pad2', ctypes.c_int32),
                ('sym', (ctypes.c_int64 * 8) * 8), ('tok_emb', ctypes.c_void_p), ('pos_emb', ctypes.c_void_p),
                ('in_proj_w', ctypes.c_void_p * 4), ('in_proj_b', ctypes.c_void_p * 4), ('out_proj_w', ctypes.c_void_p * 4),
                ('out_proj_b', ctypes.c_void_p * 4), ('lin1_w', ctypes.c_void_p * 4), ('lin1_b', ctypes.c_void_p * 4),
                ('lin2_w', ctypes.c_void_p * 4), ('lin2_b', ctypes.c_void_p * 4), ('ln1_w', ctypes.c_void_p * 4),
                ('ln1_b', ctypes.c_void_p * 4), ('ln2_w', ctypes.c_void_p * 4), ('ln2_b', ctypes.c_void_p * 4),
                ('dec_w', ctypes.c_void_p), ('dec_b', ctypes.c_void_p), ('cont_mask', ctypes.c_void_p),
                ('memo_size', ctypes.c_int64), ('ln_eps', ctypes.c_double)]


# entry points added by later kernel families register themselves here (name -> (restype, argtypes))
OPTIONAL_SIGNATURES = {}

class BrgProblem(ctypes.Structure):
    """anqs_brg_problem_t (include/anqs_b200.h): C[M][N] (+)= A^T B, colsum[M] (+)= column sums of A."""
    _fields_ = [('A', _vp), ('B', _vp), ('C', _vp), ('colsum', _vp),
                ('lda', ctypes.c_int32), ('ldb', ctypes.c_int32), ('ldc', ctypes.c_int32), ('M', ctypes.c_int32), ('N', ctypes.c_int32),
                ('reserved', ctypes.c_int32)]


_brg_workspace = {}


def batch_reduce(problems, K: int, accumulate: bool, device):
    """problems: list of (A, lda, M, B, ldb, N, C, ldc, colsum or None) with device addresses (ints).  One call of
    anqs_batch_reduce_gemm on the current stream; the workspace is kept per device and grows on demand."""
    import torch as pt
    arr = (BrgProblem * len(problems))()
    for i, (A, lda, M, B, ldb, N, C, ldc, colsum) in enumerate(problems):
        arr[i] = BrgProblem(A, B, C, colsum, lda, ldb, ldc, M, N, 0)
    need = lib().anqs_batch_reduce_workspace(arr, len(problems), K)
    if need < 0:
        raise RuntimeError('anqs_batch_reduce_workspace: bad problem description')
    ws = _brg_workspace.get(device)
    if ws is None or ws.numel() * 8 < need:
        ws = _brg_workspace[device] = pt.empty((need + 7) // 8, dtype=pt.float64, device=device)
    check(lib().anqs_batch_reduce_gemm(arr, len(problems), K, int(bool(accumulate)), ctypes.c_void_p(ws.data_ptr()), ws.numel() * 8,
                                       stream_ptr(device)))


_lib = None


def declared_symbols():
    return sorted(list(_SIGNATURES) + list(OPTIONAL_SIGNATURES))


def lib():
    """Loads libanqs_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: the CUDA library is not built. Run `make -C anqs_quantum_chemistry_b200/csrc` '
                f'(or __graft_entry__.build()). There is no CPU fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGNATURES.items()) + list(OPTIONAL_SIGNATURES.items()):
            fn = getattr(handle, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if handle.anqs_abi_version() != 1:
            raise RuntimeError('libanqs_b200.so ABI version mismatch')
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().anqs_last_error().decode('utf-8', 'replace'))


def stream_ptr(device=None):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def dptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return _vp(0)
    if not t.is_cuda:
        raise RuntimeError('expected a CUDA tensor: the anqs_b200 kernels have no CPU path')
    if not t.is_contiguous():
        raise RuntimeError('expected a contiguous tensor')
    return _vp(t.data_ptr())


def require_cuda(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError(f'device {device} is not a CUDA device: the anqs_b200 kernels have no CPU path')
    if not torch.cuda.is_available():
        raise RuntimeError('CUDA is not available: the anqs_b200 kernels have no CPU path')
    return device
