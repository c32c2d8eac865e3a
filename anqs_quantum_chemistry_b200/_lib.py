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
    'anqs_energy_stats_workspace': (ctypes.c_size_t, [_c_i64]),
    'anqs_energy_stats': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp]),
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
    'anqs_pair_join_workspace': (ctypes.c_size_t, [_c_i64, _c_i64]),
    'anqs_local_energy_pair_join': (_c_int, [_vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp]),
    'anqs_sort_workspace': (ctypes.c_size_t, [_c_i64]),
    'anqs_sort_pairs_u64': (_c_int, [_vp, _vp, _vp, _vp, _c_i64, _c_int, _c_int, _c_int, ctypes.c_uint64, _vp, _vp]),
    'anqs_unique_workspace': (ctypes.c_size_t, [_c_i64]),
    'anqs_unique_i64': (_c_int, [_vp, _c_i64, _c_int, _vp, _vp, _vp, _vp, _vp]),
    'anqs_topk_workspace': (ctypes.c_size_t, [_c_i64, _c_i64]),
    'anqs_topk_f64': (_c_int, [_vp, _c_i64, _c_i64, _c_int, _vp, _vp, _vp, _vp]),
    'anqs_local_energy_sample_aware': (_c_int, [_vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp, _vp]),
    'anqs_local_energy_sample_aware_variant': (_c_int, [_vp, _vp, _vp, _c_i64, _c_i64, _c_i64, _vp, _c_i64, _c_int, _c_int, _vp, _c_int, _vp]),
    'anqs_accumulate_rows': (_c_int, [_vp, _c_i64, _vp, _vp, _c_int, _vp, _vp, _vp, _c_int, _vp]),
    'anqs_made_log_psi': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    'anqs_made_cond_log_abs': (_c_int, [_vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_made_backward_chain': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_made_backward_chain_abs': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_made_phase_output_workspace': (_c_i64, [_vp]),
    'anqs_made_phase_output_grad': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _c_int, _vp, _vp, _vp, _c_i64, _vp]),
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
    'anqs_transformer_backward_workspace': (_c_i64, [_vp, _c_i64]),
    'anqs_transformer_log_psi_saving': (_c_int, [_vp, _vp, _c_i64, _vp, _vp, _c_i64, _vp]),
    'anqs_transformer_backward': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp, _c_i64, _c_int, _c_int, _vp]),
    'anqs_transformer_tc_packed_bytes': (ctypes.c_size_t, [_vp]),
    'anqs_transformer_tc_pack': (_c_int, [_vp, _vp, _vp]),
    'anqs_transformer_log_psi_tc': (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    'anqs_transformer_cond_log_abs_tc': (_c_int, [_vp, _vp, _c_int, _vp, _c_i64, _vp, _vp]),
    'anqs_sampler_split_level': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _c_i64, _c_i64, _c_int, _c_int, ctypes.c_uint64,
                                          _c_i64, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_emit_children': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_emit_children_capped': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    'anqs_sampler_gumbel_select': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'anqs_sampler_gumbel_select_masked': (_c_int, [_vp, _vp, _c_i64, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp,
                                                   _vp, _vp]),
    'anqs_sampler_gumbel_level': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _c_int, ctypes.c_uint64,
                                           _c_i64, _vp, _vp, _vp, _vp]),
    'anqs_sampler_gumbel_level_keyed': (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _vp, _c_i64, _c_i64, _c_int, ctypes.c_uint64,
                                                 _c_i64, _vp, _vp, _vp, _vp, _vp]),
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


class TransformerGrads(ctypes.Structure):
    """anqs_transformer_grads_t (include/anqs_b200.h): destinations of the parameter gradients."""
    _fields_ = [('tok_emb', ctypes.c_void_p), ('pos_emb', ctypes.c_void_p),
                ('in_proj_w', ctypes.c_void_p * 4), ('in_proj_b', ctypes.c_void_p * 4), ('out_proj_w', ctypes.c_void_p * 4),
                ('out_proj_b', ctypes.c_void_p * 4), ('lin1_w', ctypes.c_void_p * 4), ('lin1_b', ctypes.c_void_p * 4),
                ('lin2_w', ctypes.c_void_p * 4), ('lin2_b', ctypes.c_void_p * 4), ('ln1_w', ctypes.c_void_p * 4),
                ('ln1_b', ctypes.c_void_p * 4), ('ln2_w', ctypes.c_void_p * 4), ('ln2_b', ctypes.c_void_p * 4),
                ('dec_w', ctypes.c_void_p), ('dec_b', ctypes.c_void_p)]


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
    """The current stream of `device` as a void*.  Every launch of the library goes to the CURRENT CUDA device (grid sizes come
    from its SM count, handles and workspaces live on it), so a tensor on another device is refused here with a clear message
    instead of an 'invalid resource handle' from the driver: one process (or one `with torch.cuda.device(...)` block) per GPU."""
    if device is not None:
        device = torch.device(device)
        if device.type == 'cuda' and device.index is not None and device.index != torch.cuda.current_device():
            raise RuntimeError(f'anqs_b200: tensors live on {device} but the current CUDA device is cuda:{torch.cuda.current_device()}; '
                               f'call torch.cuda.set_device({device.index}) or wrap the call in `with torch.cuda.device({device.index}):`')
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


def _workspace(nbytes, device):
    return torch.empty(max(1, (int(nbytes) + 7) // 8), dtype=torch.int64, device=device)


def sort_pairs(keys: torch.Tensor, vals: torch.Tensor = None, begin_bit: int = 0, end_bit: int = 64, key_kind: int = 0, xor_mask: int = 0):
    """(sorted keys, payloads) through anqs_sort_pairs_u64; keys int64 or float64 [n] on a CUDA device, vals int64 [n] or None
    (payload = position, i.e. the permutation)."""
    dev = require_cuda(keys.device)
    keys = keys.contiguous()
    n = keys.shape[0]
    out_k, out_v = torch.empty_like(keys), torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return out_k, out_v
    work = _workspace(lib().anqs_sort_workspace(n), dev)
    check(lib().anqs_sort_pairs_u64(dptr(keys), dptr(vals.contiguous() if vals is not None else None), dptr(out_k), dptr(out_v), n,
                                    int(begin_bit), int(end_bit), int(key_kind), ctypes.c_uint64(xor_mask & 0xFFFFFFFFFFFFFFFF), dptr(work),
                                    stream_ptr(dev)))
    return out_k, out_v


def unique_i64(x: torch.Tensor, end_bit: int = 64, return_inverse: bool = True):
    """Sorted (signed) unique values of an int64 vector and the inverse map (anqs_unique_i64); one host read for the count."""
    dev = require_cuda(x.device)
    x = x.contiguous().view(-1)
    n = x.shape[0]
    unq = torch.empty(n, dtype=torch.int64, device=dev)
    inv = torch.empty(n, dtype=torch.int64, device=dev) if return_inverse else None
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    if n > 0:
        work = _workspace(lib().anqs_unique_workspace(n), dev)
        check(lib().anqs_unique_i64(dptr(x), n, int(end_bit), dptr(unq), dptr(inv), dptr(cnt), dptr(work), stream_ptr(dev)))
    return unq[:int(cnt.item())], inv


def topk_f64(vals: torch.Tensor, k: int, sorted: bool = True):
    """The k largest entries of a float64 vector through anqs_topk_f64: (values, positions), as the first k rows of a stable
    descending sort (sorted=True) or the same set in position order (sorted=False)."""
    dev = require_cuda(vals.device)
    vals = vals.contiguous().view(-1)
    n = vals.shape[0]
    top_v = torch.empty(k, dtype=torch.float64, device=dev)
    top_i = torch.empty(k, dtype=torch.int64, device=dev)
    if k > 0:
        work = _workspace(lib().anqs_topk_workspace(n, k), dev)
        check(lib().anqs_topk_f64(dptr(vals), n, int(k), int(bool(sorted)), dptr(top_v), dptr(top_i), dptr(work), stream_ptr(dev)))
    return top_v, top_i


def require_cuda(device):
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError(f'device {device} is not a CUDA device: the anqs_b200 kernels have no CPU path')
    if not torch.cuda.is_available():
        raise RuntimeError('CUDA is not available: the anqs_b200 kernels have no CPU path')
    return device
