"""LogAbsPhaseANQS drop-in in MADE mode (reference: nqs/nqs/stochastic/ansatzes/anqs/{abstract_anqs,
log_abs_phase_anqs,mlp}.py).

Same constructor keywords, parameter names (`log_abs_subnet.layers.{i}.{weight,bias}`,
`phase_subnet.layers.{i}.{weight,bias}`) and initialisation order as the reference, so `state_dict`s interchange and
`pt.manual_seed(s)` followed by construction gives the reference's initial weights.  Method surface used by the
reference's callers (SMP:51-101, EXP:525-545, CLE, SR, PG): amplitude, log_psi, sample_stats, sample_indices_gumbel,
cond_log_abs, cat_grad, param_num, clip_grad_norm, compute_cat_log_jac, sort_base_idx, spin_flip_base_idx.

Compute: forward passes run in the sm_100a kernels of libanqs_b200.so (k3_made.cu, k4_sampler.cu); the backward
pass of log_psi is a hand-derived chain (no torch autograd graph through the network) whose GEMMs are plain
library fp64 matmuls on the activations saved by the forward kernel.  There is no CPU path.
"""
import ctypes
import math
from typing import Tuple

import numpy as np
import torch as pt
from torch import nn

from . import _lib
from .abstract_hilbert_space_object import AbstractHilbertSpaceObject
from .constants import BASE_REAL_TYPE, BASE_COMPLEX_TYPE, NEGINF
from .masker import LocallyDecomposableMasker
from .qubit_grouping import QubitGrouping, QubitGroupingConfig
from .sampler import AutoregressiveSamplerMixin, ParameterVectorMixin

LOCAL_SAMPLING_STRATEGIES = ('DU', 'MU')


class LocalSamplingConfig:
    """ANQS:20-50."""

    def __init__(self, *args, pattern_type: str = 'uniform', strategy: str = 'MU', masking_depth: int = 0, **kwargs):
        if args or kwargs:
            raise TypeError(f'LocalSamplingConfig: unexpected arguments {args} {sorted(kwargs)}')
        self.pattern_type, self.strategy, self.masking_depth = pattern_type, strategy, masking_depth

    def create_local_sampling_pattern(self, qudit_num: int = None):
        assert self.pattern_type == 'uniform'
        assert self.strategy in LOCAL_SAMPLING_STRATEGIES
        assert 0 <= self.masking_depth <= qudit_num, f'masking_depth {self.masking_depth} outside [0, {qudit_num}]'
        return (self.strategy,) * (qudit_num - self.masking_depth) + ('DU',) * self.masking_depth


class _PatternConfig:
    """PatternConfig of the reference (infrastructure/nested_data.py:97-125): a per-layer pattern; only 'uniform' exists there
    apart from ActivationConfig's 'sanqs_paper'."""
    UNIFORM_PATTERN_FIELD = None

    def __init__(self, *args, pattern_type: str = 'uniform', **kwargs):
        if args or kwargs:
            raise TypeError(f'{type(self).__name__}: unexpected arguments {args} {sorted(kwargs)}')
        if pattern_type != 'uniform':
            raise NotImplementedError(f"{type(self).__name__}: pattern_type '{pattern_type}' is not implemented by the sm_100a kernels "
                                      f"(uniform layers only)")
        self.pattern_type = pattern_type

    def create_pattern(self, depth):
        return (getattr(self, self.UNIFORM_PATTERN_FIELD),) * depth


class WidthConfig(_PatternConfig):
    """MLP:13-23."""
    UNIFORM_PATTERN_FIELD = 'width'

    def __init__(self, *args, width: int = 64, **kwargs):
        self.width = width
        super().__init__(*args, **kwargs)


class BiasConfig(_PatternConfig):
    """MLP:26-36."""
    UNIFORM_PATTERN_FIELD = 'use_bias'

    def __init__(self, *args, use_bias: bool = True, **kwargs):
        self.use_bias = use_bias
        super().__init__(*args, **kwargs)


class ActivationConfig(_PatternConfig):
    """MLP:49-70."""
    UNIFORM_PATTERN_FIELD = 'activation'

    def __init__(self, *args, activation=nn.Tanh, **kwargs):
        self.activation = activation
        super().__init__(*args, **kwargs)


class MLPConfig:
    """MLP:83-99, same keywords: depth, width_config, use_res, bias_config, activation_config, activate_last_layer.  What the
    kernels do not implement raises NotImplementedError instead of silently building a different network: hidden width other
    than 64 (made_common.cuh MD_W), non-uniform patterns, an activation other than tanh, an activated last layer.  Unknown
    keywords raise TypeError.  `width=` / `use_bias=` / `activation=` are accepted as shorthands for the three sub-configs."""

    def __init__(self, *args, depth: int = 2, width_config: WidthConfig = None, use_res: bool = True, bias_config: BiasConfig = None,
                 activation_config: ActivationConfig = None, activate_last_layer: bool = False,
                 width: int = None, use_bias: bool = None, activation=None, **kwargs):
        if args or kwargs:
            raise TypeError(f'MLPConfig: unexpected arguments {args} {sorted(kwargs)}')
        for name, short, full in (('width', width, width_config), ('use_bias', use_bias, bias_config), ('activation', activation, activation_config)):
            if short is not None and full is not None:
                raise TypeError(f'MLPConfig: give either {name}= or its config object, not both')
        self.depth = depth
        self.width_config = width_config if width_config is not None else WidthConfig(**({} if width is None else {'width': width}))
        self.use_res = use_res
        self.bias_config = bias_config if bias_config is not None else BiasConfig(**({} if use_bias is None else {'use_bias': use_bias}))
        self.activation_config = (activation_config if activation_config is not None
                                  else ActivationConfig(**({} if activation is None else {'activation': activation})))
        self.activate_last_layer = activate_last_layer
        if not 1 <= self.depth <= 4:
            raise NotImplementedError('MLPConfig: depth must be in [1, 4] (anqs_made_desc_t holds five layers)')
        if self.width != 64:
            raise NotImplementedError(f'MLPConfig: hidden width {self.width} is not implemented (the sm_100a kernels are built for width 64, '
                                      f'the reference default)')
        if self.activation is not nn.Tanh:
            raise NotImplementedError('MLPConfig: only nn.Tanh hidden activations are implemented')
        if self.activate_last_layer:
            raise NotImplementedError('MLPConfig: activate_last_layer=True is not implemented (identity output layer only)')

    # flat views used by the modules
    width = property(lambda self: self.width_config.width)
    use_bias = property(lambda self: self.bias_config.use_bias)
    activation = property(lambda self: self.activation_config.activation)


class ANQSConfig:
    """ANQS:68-109, same keywords and the same defaults: de_mode='NADE' (one MLP pair per qudit, ANQS:90); 'MADE' (one masked
    network, the mode BASELINE.json's configurations name) has to be asked for, as in the reference."""
    ALLOWED_DE_MODES = ('MADE', 'NADE')

    def __init__(self, *args, dtype=BASE_REAL_TYPE, de_mode: str = 'NADE', qubit_grouping_config: QubitGroupingConfig = None,
                 local_sampling_config: LocalSamplingConfig = None, subtract_mean: bool = True,
                 main_subnet_config: MLPConfig = None, aux_subnet_config: MLPConfig = None,
                 use_sign_structure: bool = False, spin_flip_symmetry_config=None, **kwargs):
        if args or kwargs:
            raise TypeError(f'ANQSConfig: unexpected arguments {args} {sorted(kwargs)}')
        assert de_mode in self.ALLOWED_DE_MODES
        if use_sign_structure:
            raise NotImplementedError('ANQSConfig: use_sign_structure=True is not implemented')
        if spin_flip_symmetry_config is not None and (getattr(spin_flip_symmetry_config, 'abs', False) or getattr(spin_flip_symmetry_config, 'phase', False)):
            raise NotImplementedError('ANQSConfig: spin-flip symmetrisation is broken in the reference (ANQS:160-190) and not implemented')
        self.dtype = dtype
        self.de_mode = de_mode
        self.qubit_grouping_config = qubit_grouping_config if qubit_grouping_config is not None else QubitGroupingConfig()
        self.local_sampling_config = local_sampling_config if local_sampling_config is not None else LocalSamplingConfig()
        self.subtract_mean = subtract_mean
        self.main_subnet_config = main_subnet_config if main_subnet_config is not None else MLPConfig()
        self.aux_subnet_config = aux_subnet_config if aux_subnet_config is not None else MLPConfig()
        self.use_sign_structure = use_sign_structure
        self.spin_flip_symmetry_config = spin_flip_symmetry_config


class MLP(nn.Module):
    """MLP of the reference (MLP:102-246): parameters (and, in MADE form, the causal masks) only; the forward pass lives in
    k3_made.cu / k3_nade.cu.  is_made=True: one masked network for all qudits; is_made=False: the plain network of one qudit
    in NADE mode (in_num inputs, out_num outputs)."""

    def __init__(self, in_num: int = None, is_made: bool = True, qubit_grouping: QubitGrouping = None, out_num: int = None,
                 dtype=BASE_REAL_TYPE, is_out_complex: bool = False, config: MLPConfig = None):
        super().__init__()
        assert dtype == BASE_REAL_TYPE and not is_out_complex
        self.config = config if config is not None else MLPConfig()
        cfg = self.config
        assert cfg.activation is nn.Tanh and not cfg.activate_last_layer, 'the kernels implement tanh hidden / identity output'
        self.in_num, self.depth, self.dtype, self.is_made = in_num, cfg.depth, dtype, is_made
        width = (cfg.width,) * cfg.depth
        in_nums = (in_num,) + width
        if not is_made:
            self.out_num, self.val_per_out = out_num, 1
            self.layers = nn.ModuleList([nn.Linear(in_nums[l], (width + (out_num,))[l], bias=cfg.use_bias, dtype=dtype)
                                         for l in range(cfg.depth + 1)])
            self.made_masks = ()
            return
        assert qubit_grouping is not None
        self.out_num = qubit_grouping.qudit_num
        self.val_per_out = max(qubit_grouping.qudit_dims_host)
        out_nums = width + (self.out_num * self.val_per_out,)
        self.layers = nn.ModuleList([nn.Linear(in_nums[l], out_nums[l], bias=cfg.use_bias, dtype=dtype)
                                     for l in range(cfg.depth + 1)])
        # causal masks (MLP:170-203)
        allowed = []
        for l in range(cfg.depth):
            row = []
            for g in range(self.out_num):
                row += [g] * (width[l] // self.out_num + 1 * ((self.out_num - g - 1) < (width[l] % self.out_num)))
            allowed.append(row)
        allowed = pt.tensor(allowed)
        ends = qubit_grouping.qudit_ends
        start_connect = []
        for g in range(self.out_num):
            start_connect += [g] * (ends[g] - (ends[g - 1] if g > 0 else 0))
        start_connect = pt.tensor(start_connect)
        start_mask = pt.greater(allowed[0].unsqueeze(-1), start_connect.unsqueeze(0)).type(dtype)
        mid_masks = [pt.ge(allowed[l].unsqueeze(-1), allowed[l - 1].unsqueeze(0)).type(dtype) for l in range(1, cfg.depth)]
        end_connect = pt.arange(self.out_num).unsqueeze(-1).tile((1, self.val_per_out)).reshape(-1)
        end_mask = pt.ge(end_connect.unsqueeze(-1), allowed[-1].unsqueeze(0)).type(dtype)
        self.made_masks = (start_mask,) + tuple(mid_masks) + (end_mask,)

    def apply_made_masks_(self):
        """MLP:230-233: the reference overwrites weight.data with weight.data * mask on every forward."""
        with pt.no_grad():
            for layer, mask in zip(self.layers, self.made_masks):
                layer.weight.data.mul_(mask.to(layer.weight.device))


_MADE_BWD_SCRATCH_BYTES = 1 << 30  # scratch of one backward launch (dY, da, x); larger batches are processed in chunks


class _MadeLogPsi(pt.autograd.Function):
    """log psi(x) of a batch of packed configurations; backward by hand from the saved activations."""

    @staticmethod
    def forward(ctx, wf, idx, *params):
        need_grad = any(ctx.needs_input_grad[2:])  # all False under no_grad (grad mode is always off inside forward)
        log_psi, saved = wf._launch_log_psi(idx, save=need_grad)
        ctx.wf, ctx.saved, ctx.idx = wf, saved, idx
        ctx.weights = [p.detach() for p in params]
        return log_psi

    @staticmethod
    def backward(ctx, grad_out):
        """The per-sample chain runs in made_backward_kernel (k3_made_bwd.cu), the reductions over the batch - grad W_out =
        dY^T h_last, grad W_l = da_l^T (h_{l-1} | x) and the bias sums of both sub-networks - in batch_reduce_gemm_kernel
        (k3_batch_reduce.cu)."""
        wf, idx = ctx.wf, ctx.idx
        save_h, save_p = ctx.saved                      # [2, depth, B, width], [B, Q, DM]
        B, Q, DM, depth, n = idx.shape[0], wf.qudit_num, wf.max_qudit_dim, wf.depth, wf.qubit_num
        dev = idx.device
        g = grad_out.to(pt.complex128).contiguous()
        desc = wf._descriptor()
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        QD, width = Q * DM, wf.width
        chunk = max(1, min(B, _MADE_BWD_SCRATCH_BYTES // (8 * QD + 16 * depth * width + 8 * n)))
        f64 = dict(dtype=pt.float64, device=dev)
        alloc = pt.zeros if B == 0 else pt.empty
        gW_out, gb_out = alloc((2, QD, width), **f64), alloc((2, QD), **f64)
        gW0, gb_h = alloc((2, width, n), **f64), alloc((2, depth, width), **f64)
        gWm = alloc((2, depth - 1, width, width), **f64) if depth > 1 else None
        po_work = wf._phase_output_workspace(desc) if B > 0 else None
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            m = hi - lo
            dY = pt.empty((m, QD), **f64)                  # log-abs network only: the phase network's is one entry per qudit
            da = pt.empty((2, depth, m, width), **f64)
            x = pt.empty((m, n), **f64)
            h = save_h if m == B else save_h[:, :, lo:hi].contiguous()
            p = save_p if m == B else save_p[lo:hi]
            gl = pt.view_as_real(g[lo:hi])
            _lib.check(lib.anqs_made_backward_chain_abs(ctypes.byref(desc), _lib.dptr(idx[lo:hi]), m, _lib.dptr(gl),
                                                        _lib.dptr(h), _lib.dptr(p), _lib.dptr(dY), _lib.dptr(da), _lib.dptr(x), sp))
            # phase network, output layer: a row scatter (k3_made_bwd.cu: made_phase_output_kernel)
            _lib.check(lib.anqs_made_phase_output_grad(ctypes.byref(desc), _lib.dptr(idx[lo:hi]), m, _lib.dptr(gl), _lib.dptr(h[1, depth - 1]),
                                                       int(lo > 0), _lib.dptr(gW_out[1]), _lib.dptr(gb_out[1]), _lib.dptr(po_work),
                                                       po_work.numel() * 8, sp))
            problems = []   # k3_batch_reduce.cu: every other batch reduction of both sub-networks in one launch pair
            problems.append((dY.data_ptr(), QD, QD, h[0, depth - 1].data_ptr(), width, width, gW_out[0].data_ptr(), width, gb_out[0].data_ptr()))
            for net in range(2):
                problems.append((da[net, 0].data_ptr(), width, width, x.data_ptr(), n, n, gW0[net].data_ptr(), n, gb_h[net, 0].data_ptr()))
                for l in range(1, depth):
                    problems.append((da[net, l].data_ptr(), width, width, h[net, l - 1].data_ptr(), width, width,
                                     gWm[net, l - 1].data_ptr(), width, gb_h[net, l].data_ptr()))
            _lib.batch_reduce(problems, m, lo > 0, dev)
        n_layer = depth + 1
        per_net = 2 * n_layer if wf.use_bias else n_layer
        grads = [None] * (2 * per_net)
        for net in range(2):
            for l in range(n_layer):
                gw = gW0[net] if l == 0 else (gW_out[net] if l == depth else gWm[net, l - 1])
                if wf.use_bias:
                    grads[net * per_net + 2 * l] = gw
                    grads[net * per_net + 2 * l + 1] = gb_out[net] if l == depth else gb_h[net, l]
                else:
                    grads[net * per_net + l] = gw
        return (None, None) + tuple(grads)


class _NadeLogPsi(pt.autograd.Function):
    """NADE-mode log psi; backward by hand from the saved activations, one small chain per (sub-network, qudit)."""

    @staticmethod
    def forward(ctx, wf, idx, *params):
        need_grad = any(ctx.needs_input_grad[2:])
        log_psi, saved = wf._launch_log_psi(idx, save=need_grad)
        ctx.wf, ctx.saved, ctx.idx = wf, saved, idx
        ctx.weights = [p.detach() for p in params]
        return log_psi

    @staticmethod
    def backward(ctx, grad_out):
        """The per-sample chains of all 2 Q MLPs run in nade_backward_kernel (k3_made_bwd.cu), the reductions over the batch
        of every (sub-network, qudit, layer) in batch_reduce_gemm_kernel (k3_batch_reduce.cu)."""
        wf, idx = ctx.wf, ctx.idx
        save_h, save_p = ctx.saved                      # [2, Q, depth, B, width], [B, Q, DM]
        B, Q, DM, depth, n, width = idx.shape[0], wf.qudit_num, wf.max_qudit_dim, wf.depth, wf.qubit_num, wf.width
        dev = idx.device
        g = grad_out.to(pt.complex128).contiguous()
        desc = wf._descriptor()
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        QD = Q * DM
        chunk = max(1, min(B, _MADE_BWD_SCRATCH_BYTES // (32 * QD + 16 * Q * depth * width + 8 * n)))
        f64 = dict(dtype=pt.float64, device=dev)
        alloc = pt.zeros if B == 0 else pt.empty
        gW_out, gb_out = alloc((2, Q, DM, width), **f64), alloc((2, Q, DM), **f64)
        gW0, gb_h = alloc((2, Q, width, n), **f64), alloc((2, Q, depth, width), **f64)
        gWm = alloc((2, Q, depth - 1, width, width), **f64) if depth > 1 else None
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            m = hi - lo
            dY = pt.empty((2, m, QD), **f64)
            da = pt.empty((2, Q, depth, m, width), **f64)
            x = pt.empty((m, n), **f64)
            h = save_h if m == B else save_h[:, :, :, lo:hi].contiguous()
            p = save_p if m == B else save_p[lo:hi]
            _lib.check(lib.anqs_nade_backward_chain(ctypes.byref(desc), _lib.dptr(idx[lo:hi]), m, _lib.dptr(pt.view_as_real(g[lo:hi])),
                                                    _lib.dptr(h), _lib.dptr(p), _lib.dptr(dY), _lib.dptr(da), _lib.dptr(x), sp))
            problems = []
            for net in range(2):   # k3_batch_reduce.cu: every (sub-network, qudit, layer) reduction in one call
                for q in range(Q):
                    problems.append((dY[net].data_ptr() + 8 * q * DM, QD, DM, h[net, q, depth - 1].data_ptr(), width, width,
                                     gW_out[net, q].data_ptr(), width, gb_out[net, q].data_ptr()))
                    problems.append((da[net, q, 0].data_ptr(), width, width, x.data_ptr(), n, n, gW0[net, q].data_ptr(), n,
                                     gb_h[net, q, 0].data_ptr()))
                    for l in range(1, depth):
                        problems.append((da[net, q, l].data_ptr(), width, width, h[net, q, l - 1].data_ptr(), width, width,
                                         gWm[net, q, l - 1].data_ptr(), width, gb_h[net, q, l].data_ptr()))
            _lib.batch_reduce(problems, m, lo > 0, dev)
        n_layer = depth + 1
        per_mlp = 2 * n_layer if wf.use_bias else n_layer
        grads = [None] * len(ctx.weights)
        for net in range(2):
            for q in range(Q):
                base = (net * Q + q) * per_mlp
                start, D = wf.qudit_starts[q], wf.qubit_grouping.qudit_dims_host[q]
                for l in range(n_layer):
                    if l == 0:   # LAP:26: the first qudit's network sees one constant-zero input
                        gw = gW0[net, q, :, :start] if start > 0 else pt.zeros((width, 1), dtype=pt.float64, device=dev)
                    elif l == depth:
                        gw = gW_out[net, q, :D]
                    else:
                        gw = gWm[net, q, l - 1]
                    if l == depth and depth == 0:
                        gw = gW_out[net, q, :D]
                    if wf.use_bias:
                        grads[base + 2 * l] = gw
                        grads[base + 2 * l + 1] = gb_out[net, q, :D] if l == depth else gb_h[net, q, l]
                    else:
                        grads[base + l] = gw
        return (None, None) + tuple(grads)


class LogAbsPhaseANQS(AutoregressiveSamplerMixin, ParameterVectorMixin, AbstractHilbertSpaceObject, nn.Module):
    def __init__(self, *args, config: ANQSConfig = None, masker: LocallyDecomposableMasker = None, **kwargs):
        AbstractHilbertSpaceObject.__init__(self, *args, **kwargs)
        nn.Module.__init__(self)
        self.config = config if config is not None else ANQSConfig()
        assert self.config.de_mode in ANQSConfig.ALLOWED_DE_MODES
        assert self.config.dtype == BASE_REAL_TYPE
        assert not self.config.use_sign_structure
        self.dtype, self.de_mode = self.config.dtype, self.config.de_mode
        self.masker = masker
        self.qubit_grouping_config = self.config.qubit_grouping_config
        self.qubit_grouping = QubitGrouping.create(hs=self.hilbert_space, config=self.qubit_grouping_config, masker=masker)
        self.max_qudit_dim = max(self.qubit_grouping.qudit_dims_host)
        self.local_sampling_config = self.config.local_sampling_config
        self.local_sampling_pattern = self.local_sampling_config.create_local_sampling_pattern(qudit_num=self.qudit_num)
        main, aux = self.config.main_subnet_config, self.config.aux_subnet_config
        assert (main.depth, main.width, main.use_res, main.use_bias) == (aux.depth, aux.width, aux.use_res, aux.use_bias), \
            'the kernel evaluates both sub-networks with one shape'
        self.depth, self.width, self.use_res, self.use_bias = main.depth, main.width, main.use_res, main.use_bias
        # construction order = reference order (LAP:43-56), so the global torch RNG yields the same initial weights
        if self.de_mode == 'MADE':
            self.log_abs_subnet = MLP(in_num=self.qubit_num, is_made=True, qubit_grouping=self.qubit_grouping, dtype=self.dtype, config=main)
            self.phase_subnet = MLP(in_num=self.qubit_num, is_made=True, qubit_grouping=self.qubit_grouping, dtype=self.dtype, config=aux)
        else:  # LAP:24-42: one plain MLP per qudit, all log-abs networks first, then all phase networks
            qg = self.qubit_grouping
            def per_qudit(cfg):
                return nn.ModuleList([MLP(in_num=qg.qudit_ends[q - 1] if q != 0 else 1, is_made=False, qubit_grouping=qg,
                                          out_num=qg.qudit_dims_host[q], dtype=self.dtype, config=cfg) for q in range(qg.qudit_num)])
            self.log_abs_subnet = per_qudit(main)
            self.phase_subnet = per_qudit(aux)
        self._ptr_table = None
        self._ptr_key = None
        self._desc_cache = None        # (parameter pointers, descriptor): rebuilt only when a parameter is re-allocated
        self._param_list = None
        self.to(self.device)
        self._param_num = None
        self._masked_key = None        # parameter versions right after the MADE masks were last applied
        self._packed_tc = None         # weights in the tensor cores' operand layout (k3_made_tc.cu)
        self._packed_key = None
        self.inference_precision = 'fp64'   # 'tf32': no-grad amplitudes and the samplers' conditionals run on tcgen05
        self._init_sampler()

    # ---- shapes ------------------------------------------------------------------------------------------------
    qudit_num = property(lambda self: self.qubit_grouping.qudit_num)
    qudit_dims = property(lambda self: self.qubit_grouping.qudit_dims)
    qudit_starts = property(lambda self: self.qubit_grouping.qudit_starts)
    qudit_ends = property(lambda self: self.qubit_grouping.qudit_ends)

    # ---- kernel plumbing ---------------------------------------------------------------------------------------
    def set_inference_precision(self, precision: str):
        """'fp64' (default; the reference's precision, every path) or 'tf32' (tcgen05 tensor cores for evaluations that do
        not need gradients: amplitudes of non-sampled configurations, the samplers' conditional probabilities)."""
        assert precision in ('fp64', 'tf32')
        self.inference_precision = precision

    def invalidate_caches(self):
        """Forget everything derived from the parameter VALUES: the MADE re-masking mark, the tensor-core packed weights, the
        kernel descriptors.  The caches are keyed on (tensor version, data pointer), which in-place writes through `p.data`
        (`p.data.copy_()`, EMA updates, weight surgery - the reference's own idiom) do not change: call this after such a
        write.  load_state_dict() and .to() / .double() / ... call it themselves; optimiser steps and `p.copy_()` under
        no_grad bump the version and need nothing."""
        for name in ('_masked_key', '_packed_key', '_desc_cache', '_ptr_key', '_param_list'):
            if hasattr(self, name):
                setattr(self, name, None)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_caches()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_caches()
        return out

    def _params(self):
        """The parameters as a plain list: nn.Module.parameters() walks the module tree on every call (~70 us here), and this is
        asked for on every kernel launch."""
        if self._param_list is None:
            self._param_list = list(self.parameters())
        return self._param_list

    def _param_key(self):
        return tuple((p._version, p.data_ptr()) for p in self._params())

    def _nade_descriptor(self) -> _lib.NadeDesc:
        dev = _lib.require_cuda(self.device)
        # the descriptor holds device POINTERS: it stays valid while no parameter is re-allocated (in-place updates keep them)
        ptr_key = tuple(p.data_ptr() for p in self._params())
        if self._desc_cache is not None and self._desc_cache[0] == ptr_key:
            return self._desc_cache[1]
        d = _lib.NadeDesc()
        qg = self.qubit_grouping
        d.qubit_num, d.qudit_num, d.max_qudit_dim = self.qubit_num, qg.qudit_num, self.max_qudit_dim
        d.depth, d.width, d.use_res = self.depth, self.width, int(self.use_res)
        d.subtract_mean, d.sym_num = int(self.config.subtract_mean), self.masker.sym_num
        for q in range(qg.qudit_num):
            d.qudit_starts[q] = qg.qudit_starts[q]
            d.du[q] = 1 if self.local_sampling_pattern[q] == 'DU' else 0
        d.qudit_starts[qg.qudit_num] = self.qubit_num
        for s, row in enumerate(self.masker.symmetry_descriptors()):
            for j, v in enumerate(row):
                d.sym[s][j] = int(v)
        ptrs, keep = [], []
        for nets in (self.log_abs_subnet, self.phase_subnet):
            for mlp in nets:
                for layer in mlp.layers:
                    w = layer.weight.data
                    assert w.is_contiguous() and w.dtype == pt.float64 and w.device == dev
                    keep.append(w)
                    ptrs += [w.data_ptr(), layer.bias.data.data_ptr() if layer.bias is not None else 0]
        key = tuple(ptrs)
        if self._ptr_key != key:  # the table only changes when a parameter is re-allocated
            self._ptr_table = pt.tensor(ptrs, dtype=pt.int64, device=dev)
            self._ptr_key = key
        d.ptrs = self._ptr_table.data_ptr()
        d.cont_mask = qg.cont_mask_words.data_ptr()
        d.memo_size = self.masker.memo_size
        d._keep = keep
        self._desc_cache = (ptr_key, d)
        return d

    def _phase_output_workspace(self, desc) -> pt.Tensor:
        """Per-CTA partial sums of the phase network's output-layer gradient (anqs_made_phase_output_grad), kept between calls."""
        need = int(_lib.lib().anqs_made_phase_output_workspace(ctypes.byref(desc)))
        ws = getattr(self, '_po_work', None)
        if ws is None or ws.numel() * 8 < need or ws.device != self.device:
            ws = self._po_work = pt.empty((need + 7) // 8, dtype=pt.float64, device=self.device)
        return ws

    def _descriptor(self):
        if self.de_mode == 'NADE':
            return self._nade_descriptor()
        dev = _lib.require_cuda(self.device)
        if self._param_key() != self._masked_key:
            # MLP:230-233 re-masks on every forward; masking is idempotent, so it is skipped while no parameter changed
            self.log_abs_subnet.apply_made_masks_()
            self.phase_subnet.apply_made_masks_()
            self._masked_key = self._param_key()
        ptr_key = tuple(k[1] for k in self._masked_key)
        if self._desc_cache is not None and self._desc_cache[0] == ptr_key:
            return self._desc_cache[1]  # pointers unchanged: the descriptor built earlier is still right
        d = _lib.MadeDesc()
        qg = self.qubit_grouping
        d.qubit_num, d.qudit_num, d.max_qudit_dim = self.qubit_num, qg.qudit_num, self.max_qudit_dim
        d.depth, d.width, d.use_res = self.depth, self.width, int(self.use_res)
        d.subtract_mean, d.sym_num = int(self.config.subtract_mean), self.masker.sym_num
        for q in range(qg.qudit_num):
            d.qudit_starts[q] = qg.qudit_starts[q]
            d.du[q] = 1 if self.local_sampling_pattern[q] == 'DU' else 0
        d.qudit_starts[qg.qudit_num] = self.qubit_num
        for s, row in enumerate(self.masker.symmetry_descriptors()):
            for j, v in enumerate(row):
                d.sym[s][j] = int(v)
        keep = []
        for name, net in (('abs', self.log_abs_subnet), ('phase', self.phase_subnet)):
            for l, layer in enumerate(net.layers):
                w = layer.weight.data
                assert w.is_contiguous() and w.dtype == pt.float64 and w.device == dev
                getattr(d, f'w_{name}')[l] = w.data_ptr()
                getattr(d, f'b_{name}')[l] = layer.bias.data.data_ptr() if layer.bias is not None else None
                keep.append(w)
        d.cont_mask = qg.cont_mask_words.data_ptr()
        d.memo_size = self.masker.memo_size
        d._keep = keep
        self._desc_cache = (ptr_key, d)
        return d

    def _launch_log_psi(self, idx: pt.Tensor, save: bool):
        dev = _lib.require_cuda(self.device)
        assert idx.dtype == pt.int64 and idx.dim() == 1 and idx.is_contiguous() and idx.device == dev
        B = idx.shape[0]
        out = pt.empty(B, dtype=pt.complex128, device=dev)
        h_shape = (2, self.depth, B, self.width) if self.de_mode == 'MADE' else (2, self.qudit_num, self.depth, B, self.width)
        save_h = pt.empty(h_shape, dtype=pt.float64, device=dev) if save else None
        save_p = pt.empty((B, self.qudit_num, self.max_qudit_dim), dtype=pt.float64, device=dev) if save else None
        desc = self._descriptor()
        if self.de_mode == 'NADE':
            _lib.check(_lib.lib().anqs_nade_log_psi(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(out)),
                                                    _lib.dptr(save_h), _lib.dptr(save_p), _lib.stream_ptr(dev)))
            return out, (save_h, save_p)
        _lib.check(_lib.lib().anqs_made_log_psi(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(out)),
                                                _lib.dptr(save_h), _lib.dptr(save_p), _lib.stream_ptr(dev)))
        return out, (save_h, save_p)

    def _packed_weights(self, desc):
        """Weights packed for the tensor-core kernels; repacked when a parameter changed."""
        dev = self.device
        nade = self.de_mode == 'NADE'
        key = self._param_key() if nade else self._masked_key
        if self._packed_tc is None or self._packed_key != key:
            lib = _lib.lib()
            nbytes = int((lib.anqs_nade_tc_packed_bytes if nade else lib.anqs_made_tc_packed_bytes)(ctypes.byref(desc)))
            if self._packed_tc is None or self._packed_tc.numel() * 8 < nbytes:
                self._packed_tc = pt.empty((nbytes + 7) // 8, dtype=pt.int64, device=dev)
            _lib.check((lib.anqs_nade_tc_pack if nade else lib.anqs_made_tc_pack)(ctypes.byref(desc), _lib.dptr(self._packed_tc),
                                                                                  _lib.stream_ptr(dev)))
            self._packed_key = key
        return self._packed_tc

    @pt.no_grad()
    def log_psi_tc(self, base_idx: pt.Tensor) -> pt.Tensor:
        """log psi through the tcgen05 kernels (tf32 products, fp32 accumulation); no gradients."""
        dev = _lib.require_cuda(self.device)
        idx = base_idx.contiguous().view(-1)
        B = idx.shape[0]
        out = pt.empty(B, dtype=pt.complex128, device=dev)
        desc = self._descriptor()
        packed = self._packed_weights(desc)
        fn = _lib.lib().anqs_nade_log_psi_tc if self.de_mode == 'NADE' else _lib.lib().anqs_made_log_psi_tc
        _lib.check(fn(ctypes.byref(desc), _lib.dptr(packed), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(out)), _lib.stream_ptr(dev)))
        return out

    def chosen_outcomes(self, idx: pt.Tensor) -> pt.Tensor:
        """[B, Q] local outcome index of every qudit (QG:148-154) straight from the packed word."""
        starts = pt.tensor(self.qudit_starts, dtype=pt.int64, device=idx.device)
        widths = pt.tensor(self.qubit_grouping.qubits_per_qudit, dtype=pt.int64, device=idx.device)
        return (idx.view(-1, 1) >> starts) & ((1 << widths) - 1)

    # ---- reference surface ---------------------------------------------------------------------------------------
    def log_psi_of_indices(self, base_idx: pt.Tensor) -> pt.Tensor:
        idx = base_idx.contiguous().view(-1)
        if self.inference_precision == 'tf32' and not pt.is_grad_enabled():
            return self.log_psi_tc(idx)
        if self.de_mode == 'NADE':
            return _NadeLogPsi.apply(self, idx, *self._params())
        return _MadeLogPsi.apply(self, idx, *self._params())

    def log_psi(self, base_vec: pt.Tensor, just_return: bool = False) -> pt.Tensor:
        """ANQS:407-481 (argument is the unpacked bit matrix, as in the reference)."""
        return self.log_psi_of_indices(self.base_vec2base_idx(base_vec))

    def amplitude(self, base_idx: pt.Tensor) -> pt.Tensor:
        """ANQS:483-485.  The returned tensor carries the log psi it was exponentiated from (`amps.log_psi`, same autograd
        graph) so that calculations.vmc_loss can skip the exp -> log round trip of the reference."""
        lp = self.log_psi_of_indices(base_idx)
        amps = pt.exp(lp)
        amps.log_psi = lp
        return amps

    def phase(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.log_psi_of_indices(base_idx).imag

    def forward(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.amplitude(base_idx)

    @pt.no_grad()
    def cond_log_abs(self, qudit_idx: int = None, base_vec: pt.Tensor = None, return_all_if_made: bool = False,
                     mask: pt.Tensor = None, prefix_idx: pt.Tensor = None) -> pt.Tensor:
        """LAP:105-163 for one qudit: [B, max_qudit_dim] normalised conditional log|psi| (-inf where masked).
        `mask` is accepted for signature compatibility; the kernel derives it from the prefix (QG:199-213)."""
        assert not return_all_if_made
        dev = _lib.require_cuda(self.device)
        if prefix_idx is None:
            prefix_idx = self.hilbert_space.base_vec2base_idx(base_vec).view(-1) if base_vec.shape[-1] > 0 else \
                pt.zeros(base_vec.shape[0], dtype=pt.int64, device=dev)
        prefix_idx = prefix_idx.contiguous().view(-1)
        B = prefix_idx.shape[0]
        out = pt.empty((B, self.max_qudit_dim), dtype=pt.float64, device=dev)
        desc = self._descriptor()
        if self.inference_precision == 'tf32':
            packed = self._packed_weights(desc)
            fn = _lib.lib().anqs_nade_cond_log_abs_tc if self.de_mode == 'NADE' else _lib.lib().anqs_made_cond_log_abs_tc
            _lib.check(fn(ctypes.byref(desc), _lib.dptr(packed), qudit_idx, _lib.dptr(prefix_idx), B, _lib.dptr(out), _lib.stream_ptr(dev)))
        elif self.de_mode == 'NADE':
            _lib.check(_lib.lib().anqs_nade_cond_log_abs(ctypes.byref(desc), qudit_idx, _lib.dptr(prefix_idx), B, _lib.dptr(out),
                                                         _lib.stream_ptr(dev)))
        else:
            _lib.check(_lib.lib().anqs_made_cond_log_abs(ctypes.byref(desc), qudit_idx, _lib.dptr(prefix_idx), B, _lib.dptr(out),
                                                         _lib.stream_ptr(dev)))
        return out

    # ---- per-sample log-Jacobian for stochastic reconfiguration (ANQS:820-839) ----------------------------------
    @pt.no_grad()
    def compute_cat_log_jac(self, indices: pt.Tensor) -> pt.Tensor:
        """ANQS:820-839: [B, param_num] complex128, row b = d log(conj psi(x_b)) / d theta with the parameters concatenated in
        .parameters() order.  One forward launch with saved activations, one launch of the backward chain kernel with unit
        upstream gradients, then per-sample outer products - no loop over samples (the reference vmaps autograd over the
        <= 50 samples SR uses, SR:20-32)."""
        idx = indices.contiguous().view(-1)
        B, Q, depth, dev = idx.shape[0], self.qudit_num, self.depth, idx.device
        _, (save_h, save_p) = self._launch_log_psi(idx, save=True)
        zero = pt.zeros((), dtype=pt.float64, device=dev)
        blocks = []
        if self.de_mode == 'MADE':
            # the per-sample chain of both sub-networks in one launch (k3_made_bwd.cu) with unit upstream gradients; the rows of
            # the Jacobian are the per-sample outer products of its outputs with the saved activations
            QD, width, n = Q * self.max_qudit_dim, self.width, self.qubit_num
            dY = pt.empty((2, B, QD), dtype=pt.float64, device=dev)
            da = pt.empty((2, depth, B, width), dtype=pt.float64, device=dev)
            x = pt.empty((B, n), dtype=pt.float64, device=dev)
            ones = pt.ones((B, 2), dtype=pt.float64, device=dev)
            desc = self._descriptor()
            _lib.check(_lib.lib().anqs_made_backward_chain(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(ones), _lib.dptr(save_h),
                                                           _lib.dptr(save_p), _lib.dptr(dY), _lib.dptr(da), _lib.dptr(x),
                                                           _lib.stream_ptr(dev)))
            for net in range(2):
                cols = []
                for l in range(depth + 1):
                    if l == depth:
                        g_out, inp = dY[net], save_h[net, depth - 1]
                    else:
                        g_out, inp = da[net, l], (x if l == 0 else save_h[net, l - 1])
                    cols.append((g_out.unsqueeze(2) * inp.unsqueeze(1)).reshape(B, -1))
                    if self.use_bias:
                        cols.append(g_out)
                g = pt.cat(cols, dim=1)
                blocks.append(pt.complex(zero.expand_as(g), -g) if net == 1 else pt.complex(g, zero.expand_as(g)))
        else:
            # NADE: the chains of all 2 Q MLPs in one launch, then per-sample outer products in .parameters() order
            DM, width, n = self.max_qudit_dim, self.width, self.qubit_num
            dY = pt.empty((2, B, Q * DM), dtype=pt.float64, device=dev)
            da = pt.empty((2, Q, depth, B, width), dtype=pt.float64, device=dev)
            x = pt.empty((B, n), dtype=pt.float64, device=dev)
            ones = pt.ones((B, 2), dtype=pt.float64, device=dev)
            desc = self._descriptor()
            _lib.check(_lib.lib().anqs_nade_backward_chain(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(ones), _lib.dptr(save_h),
                                                           _lib.dptr(save_p), _lib.dptr(dY), _lib.dptr(da), _lib.dptr(x),
                                                           _lib.stream_ptr(dev)))
            dYq = dY.view(2, B, Q, DM)
            for net in range(2):
                cols = []
                for q in range(Q):
                    start, D = self.qudit_starts[q], self.qubit_grouping.qudit_dims_host[q]
                    for l in range(depth + 1):
                        if l == depth:
                            g_out, inp = dYq[net, :, q, :D], save_h[net, q, depth - 1]
                        elif l == 0:
                            g_out = da[net, q, 0]
                            inp = x[:, :start] if start > 0 else pt.zeros((B, 1), dtype=pt.float64, device=dev)
                        else:
                            g_out, inp = da[net, q, l], save_h[net, q, l - 1]
                        cols.append((g_out.unsqueeze(2) * inp.unsqueeze(1)).reshape(B, -1))
                        if self.use_bias:
                            cols.append(g_out)
                g = pt.cat(cols, dim=1)
                blocks.append(pt.complex(zero.expand_as(g), -g) if net == 1 else pt.complex(g, zero.expand_as(g)))
        return pt.cat(blocks, dim=1)

    @staticmethod
    def spin_flip_base_vec(base_vec):
        assert (base_vec.shape[-1] % 2) == 0
        return pt.stack((base_vec[..., 1::2], base_vec[..., ::2]), dim=-1).reshape(base_vec.shape)

    def spin_flip_base_idx(self, base_idx):
        return self.base_vec2base_idx(self.spin_flip_base_vec(self.base_idx2base_vec(base_idx)))
