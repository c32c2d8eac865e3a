"""Autoregressive transformer wave function (BASELINE config 3).

Network = the reference's TransformerMADE (nqs/nqs/stochastic/ansatzes/legacy/anqs_primitives/made/transformer_made.py:9-48)
with the same sub-module names, so `state_dict`s interchange; conditional log-amplitudes as in RealLogPsiTransformerMADE
(legacy/made/real_log_psi_transformer_made.py:42-58): the decoder's 4 numbers per position are (re, im) of outcome 0 and
of outcome 1.  The reference's legacy ANQS scaffolding around it is not constructible; here the network
plugs into the live framework instead: one qubit per qudit, continuation masks from the LocallyDecomposableMasker
(QG:199-213), masked normalisation re -= 0.5 logsumexp(2 re) (ANQS:392-405), samplers of AbstractANQS (ANQS:494-818).

Compute: amplitudes and the samplers' conditionals run in the hand-written fp64 kernel k5_transformer.cu; the gradient
with respect to the parameters in k5_transformer_bwd.cu (per-sample chain with the forward pass recomputed per tile) plus
the batch reductions of k3_batch_reduce.cu, wired into autograd by _TransformerLogPsi below.  `log_psi_torch` - the same
function through the torch module - is kept only as the fp64 cross-check of both (tests/test_gpu_transformer.py, 1e-10).
"""
import ctypes
import math

import torch as pt
from torch import nn

from . import _lib
from .abstract_hilbert_space_object import AbstractHilbertSpaceObject
from .constants import BASE_REAL_TYPE, BASE_COMPLEX_TYPE
from .masker import LocallyDecomposableMasker
from .qubit_grouping import QubitGrouping, QubitGroupingConfig
from .sampler import AutoregressiveSamplerMixin, ParameterVectorMixin


class _ElementwiseLayerNorm(nn.LayerNorm):
    """nn.LayerNorm with the same parameters and state_dict keys whose forward is written out in elementwise operations:
    torch's fused fp64 layer_norm kernel takes 1.8 ms per call on a [2e5, 64] input on B200 (RowwiseMomentsCUDAKernel<double>),
    7 ms of a 30 ms VMC iteration; the arithmetic below is the same function (biased variance, eps inside the square root)."""

    def forward(self, x: pt.Tensor) -> pt.Tensor:
        mu = x.mean(dim=-1, keepdim=True)
        xc = x - mu
        var = (xc * xc).mean(dim=-1, keepdim=True)
        return xc * pt.rsqrt(var + self.eps) * self.weight + self.bias


class TransformerMADE(nn.Module):
    """Token + positional embedding, causal post-norm encoder, linear decoder.  Construction order (and therefore the
    initial weights under a given torch seed) follows transformer_made.py:26-41."""

    def __init__(self, dim: int = None, out_dim: int = None, depth: int = None, qubit_num: int = None, head_num: int = None,
                 dtype=BASE_REAL_TYPE):
        super().__init__()
        self.dim, self.out_dim, self.depth, self.qubit_num, self.head_num, self.dtype = dim, out_dim, depth, qubit_num, head_num, dtype
        self.pos_embedding = nn.Embedding(qubit_num + 1, dim, dtype=dtype)
        self.embedding = nn.Embedding(3, dim, dtype=dtype)
        layer = nn.TransformerEncoderLayer(d_model=dim, nhead=head_num, dim_feedforward=dim, dropout=0.0, batch_first=True, dtype=dtype)
        self.transformer = nn.TransformerEncoder(encoder_layer=layer, num_layers=depth, enable_nested_tensor=False)
        for enc in self.transformer.layers:  # same parameters (names, values, order), elementwise forward
            for name in ('norm1', 'norm2'):
                old = getattr(enc, name)
                new = _ElementwiseLayerNorm(old.normalized_shape, eps=old.eps, dtype=dtype)
                new.weight, new.bias = old.weight, old.bias
                setattr(enc, name, new)
        self.decoder = nn.Linear(dim, out_dim, dtype=dtype)

    def forward(self, x: pt.Tensor) -> pt.Tensor:
        """x [B, L] of bits (L <= qubit_num) -> [B, L + 1, out_dim]; position t sees BOS and bits < t."""
        seq = pt.cat((pt.full((x.shape[0], 1), 2, dtype=pt.long, device=x.device), x.long()), dim=-1)
        length = seq.shape[-1]
        causal = pt.triu(pt.full((length, length), float('-inf'), dtype=self.dtype, device=seq.device), diagonal=1)
        h = self.embedding(seq) + self.pos_embedding(pt.arange(length, device=seq.device))
        return self.decoder(self.transformer(h, mask=causal, is_causal=True))


_BWD_SCRATCH_BYTES = 8 << 30   # workspace of the backward pass (activations + per-row gradient signals, ~18 kB per token)


class _TransformerLogPsi(pt.autograd.Function):
    """log psi of the transformer wave function with a hand-written backward (include/anqs_b200.h: anqs_transformer_backward).
    When the activations of the whole batch fit the workspace the forward pass keeps them (anqs_transformer_log_psi_saving)
    and the backward pass starts from them; otherwise the forward pass is the inference kernel and the backward pass walks
    the batch in chunks, recomputing the activations of each."""

    @staticmethod
    def forward(ctx, wf, idx, *params):
        ctx.wf, ctx.idx = wf, idx
        dev, B = idx.device, idx.shape[0]
        lib, desc = _lib.lib(), wf._descriptor()
        need = int(lib.anqs_transformer_backward_workspace(ctypes.byref(desc), B)) if B > 0 else 0
        if need < 0:
            raise RuntimeError('anqs_transformer_backward_workspace: unsupported network shape')
        ctx.saved_in = None
        if 0 < need <= _BWD_SCRATCH_BYTES:
            work = wf._bwd_workspace(need)
            out = pt.empty(B, dtype=pt.complex128, device=dev)
            _lib.check(lib.anqs_transformer_log_psi_saving(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(out)),
                                                           _lib.dptr(work), work.numel() * 8, _lib.stream_ptr(dev)))
            wf._bwd_token += 1
            ctx.saved_in = (work, wf._bwd_token)
            return out
        return wf.log_psi_kernel(idx, precision='fp64')

    @staticmethod
    def backward(ctx, grad_out):
        wf, idx = ctx.wf, ctx.idx
        dev = idx.device
        B = idx.shape[0]
        g = grad_out.to(pt.complex128).contiguous()
        names = [n for n, _ in wf.named_parameters()]
        params = wf._params()
        alloc = pt.zeros if B == 0 else pt.empty
        grads = {n: alloc(p.shape, dtype=pt.float64, device=dev) for n, p in zip(names, params)}
        desc = wf._descriptor()
        gd = _lib.TransformerGrads()
        pre = 'transformer_made.'
        gd.tok_emb, gd.pos_emb = grads[pre + 'embedding.weight'].data_ptr(), grads[pre + 'pos_embedding.weight'].data_ptr()
        gd.dec_w, gd.dec_b = grads[pre + 'decoder.weight'].data_ptr(), grads[pre + 'decoder.bias'].data_ptr()
        for l in range(wf.config.depth):
            lp = f'{pre}transformer.layers.{l}.'
            for field, name in (('in_proj_w', 'self_attn.in_proj_weight'), ('in_proj_b', 'self_attn.in_proj_bias'),
                                ('out_proj_w', 'self_attn.out_proj.weight'), ('out_proj_b', 'self_attn.out_proj.bias'),
                                ('lin1_w', 'linear1.weight'), ('lin1_b', 'linear1.bias'), ('lin2_w', 'linear2.weight'),
                                ('lin2_b', 'linear2.bias'), ('ln1_w', 'norm1.weight'), ('ln1_b', 'norm1.bias'),
                                ('ln2_w', 'norm2.weight'), ('ln2_b', 'norm2.bias')):
                getattr(gd, field)[l] = grads[lp + name].data_ptr()
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        if B > 0:
            saved = ctx.saved_in is not None and ctx.saved_in[1] == wf._bwd_token and ctx.saved_in[0] is wf._bwd_work
            if saved:   # the workspace still holds this call's activations (no other forward with gradients ran since)
                work = ctx.saved_in[0]
                _lib.check(lib.anqs_transformer_backward(ctypes.byref(desc), ctypes.byref(gd), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(g)),
                                                         _lib.dptr(work), work.numel() * 8, 1, 0, sp))
            else:
                per_sample = max(1, int(lib.anqs_transformer_backward_workspace(ctypes.byref(desc), 1024)) // 1024)
                chunk = max(1, min(B, _BWD_SCRATCH_BYTES // per_sample))
                wf._bwd_token += 1   # the workspace is about to be overwritten
                work = wf._bwd_workspace(int(lib.anqs_transformer_backward_workspace(ctypes.byref(desc), chunk)))
                for lo in range(0, B, chunk):
                    m = min(B, lo + chunk) - lo
                    _lib.check(lib.anqs_transformer_backward(ctypes.byref(desc), ctypes.byref(gd), _lib.dptr(idx[lo:lo + m]), m,
                                                             _lib.dptr(pt.view_as_real(g[lo:lo + m])), _lib.dptr(work), work.numel() * 8,
                                                             0, int(lo > 0), sp))
        return (None, None) + tuple(grads[n] for n in names)


class TransformerANQSConfig:
    def __init__(self, *args, dim: int = 64, depth: int = 2, head_num: int = 4, dtype=BASE_REAL_TYPE, **kwargs):
        self.dim, self.depth, self.head_num, self.dtype = dim, depth, head_num, dtype


class TransformerANQS(AutoregressiveSamplerMixin, ParameterVectorMixin, AbstractHilbertSpaceObject, nn.Module):
    def __init__(self, *args, config: TransformerANQSConfig = None, masker: LocallyDecomposableMasker = None, **kwargs):
        AbstractHilbertSpaceObject.__init__(self, *args, **kwargs)
        nn.Module.__init__(self)
        self.config = config if config is not None else TransformerANQSConfig()
        assert self.config.dtype == BASE_REAL_TYPE
        assert self.config.dim == 64, 'k5_transformer.cu is built for model dimension 64'
        self.dtype = self.config.dtype
        self.masker = masker
        self.qubit_grouping = QubitGrouping.create(hs=self.hilbert_space, config=QubitGroupingConfig(qubit_per_qudit=1), masker=masker)
        self.max_qudit_dim = 2
        self.transformer_made = TransformerMADE(dim=self.config.dim, out_dim=4, depth=self.config.depth, qubit_num=self.qubit_num,
                                                head_num=self.config.head_num, dtype=self.dtype)
        self.to(self.device)
        self._param_num = None
        self._param_list = None
        self._desc_cache = None          # (parameter pointers, descriptor)
        self._packed_tc = None           # parameters in the tensor cores' operand layout (k5_transformer_tc.cu)
        self._packed_key = None
        self.inference_precision = 'fp64'   # 'tf32': no-grad amplitudes and the samplers' conditionals run on tcgen05
        self._bwd_work = None            # workspace of the backward kernel, kept between iterations
        self._bwd_token = 0              # which forward call's activations the workspace holds
        self._init_sampler()

    qudit_num = property(lambda self: self.qubit_grouping.qudit_num)

    def set_inference_precision(self, precision: str):
        """'fp64' (default, parity mode) or 'tf32': no-grad log psi / amplitudes and the conditionals the samplers ask for run
        through the tensor-core kernels.  Gradients always run in fp64."""
        assert precision in ('fp64', 'tf32')
        if precision == 'tf32':
            assert self.config.head_num in (4, 8, 16), 'the tensor-core mode needs a head dimension <= 16'
        self.inference_precision = precision

    def _params(self):
        if self._param_list is None:
            self._param_list = list(self.parameters())
        return self._param_list

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

    def _packed_weights(self, desc):
        """Parameters packed for the tensor-core kernels; repacked when one of them changed."""
        key = tuple((p._version, p.data_ptr()) for p in self._params())
        if self._packed_tc is None or self._packed_key != key:
            nbytes = int(_lib.lib().anqs_transformer_tc_packed_bytes(ctypes.byref(desc)))
            if self._packed_tc is None or self._packed_tc.numel() * 8 < nbytes:
                self._packed_tc = pt.empty((nbytes + 7) // 8, dtype=pt.int64, device=self.device)
            _lib.check(_lib.lib().anqs_transformer_tc_pack(ctypes.byref(desc), _lib.dptr(self._packed_tc), _lib.stream_ptr(self.device)))
            self._packed_key = key
        return self._packed_tc

    def _bwd_workspace(self, nbytes: int) -> pt.Tensor:
        if self._bwd_work is None or self._bwd_work.numel() * 8 < nbytes:
            self._bwd_work = None
            self._bwd_work = pt.empty((nbytes + 7) // 8, dtype=pt.float64, device=self.device)
        return self._bwd_work

    # ---- kernel plumbing ---------------------------------------------------------------------------------------------------
    def _descriptor(self) -> _lib.TransformerDesc:
        dev = _lib.require_cuda(self.device)
        ptr_key = tuple(p.data_ptr() for p in self._params())
        if self._desc_cache is not None and self._desc_cache[0] == ptr_key:
            return self._desc_cache[1]  # the descriptor holds pointers: valid while no parameter is re-allocated
        net = self.transformer_made
        d = _lib.TransformerDesc()
        d.qubit_num, d.dim, d.depth, d.head_num, d.sym_num = self.qubit_num, net.dim, net.depth, net.head_num, self.masker.sym_num
        for s, row in enumerate(self.masker.symmetry_descriptors()):
            for j, v in enumerate(row):
                d.sym[s][j] = int(v)
        keep = []

        def ptr(t):
            if t is None:
                return None
            t = t.data
            assert t.is_contiguous() and t.dtype == pt.float64 and t.device == dev
            keep.append(t)
            return t.data_ptr()

        d.tok_emb, d.pos_emb = ptr(net.embedding.weight), ptr(net.pos_embedding.weight)
        for l, layer in enumerate(net.transformer.layers):
            d.in_proj_w[l], d.in_proj_b[l] = ptr(layer.self_attn.in_proj_weight), ptr(layer.self_attn.in_proj_bias)
            d.out_proj_w[l], d.out_proj_b[l] = ptr(layer.self_attn.out_proj.weight), ptr(layer.self_attn.out_proj.bias)
            d.lin1_w[l], d.lin1_b[l] = ptr(layer.linear1.weight), ptr(layer.linear1.bias)
            d.lin2_w[l], d.lin2_b[l] = ptr(layer.linear2.weight), ptr(layer.linear2.bias)
            d.ln1_w[l], d.ln1_b[l] = ptr(layer.norm1.weight), ptr(layer.norm1.bias)
            d.ln2_w[l], d.ln2_b[l] = ptr(layer.norm2.weight), ptr(layer.norm2.bias)
            d.ln_eps = float(layer.norm1.eps)
        d.dec_w, d.dec_b = ptr(net.decoder.weight), ptr(net.decoder.bias)
        d.cont_mask = self.qubit_grouping.cont_mask_words.data_ptr()
        d.memo_size = self.masker.memo_size
        d._keep = keep
        self._desc_cache = (ptr_key, d)
        return d

    @pt.no_grad()
    def log_psi_kernel(self, base_idx: pt.Tensor, precision: str = None) -> pt.Tensor:
        dev = _lib.require_cuda(self.device)
        idx = base_idx.contiguous().view(-1)
        B = idx.shape[0]
        out = pt.empty(B, dtype=pt.complex128, device=dev)
        desc = self._descriptor()
        if (precision or self.inference_precision) == 'tf32':
            packed = self._packed_weights(desc)
            _lib.check(_lib.lib().anqs_transformer_log_psi_tc(ctypes.byref(desc), _lib.dptr(packed), _lib.dptr(idx), B,
                                                              _lib.dptr(pt.view_as_real(out)), _lib.stream_ptr(dev)))
        else:
            _lib.check(_lib.lib().anqs_transformer_log_psi(ctypes.byref(desc), _lib.dptr(idx), B, _lib.dptr(pt.view_as_real(out)),
                                                           _lib.stream_ptr(dev)))
        return out

    def log_psi_torch(self, base_idx: pt.Tensor) -> pt.Tensor:
        """The same function through the torch module (differentiable)."""
        _lib.require_cuda(self.device)
        idx = base_idx.contiguous().view(-1, 1)
        bits = self.base_idx2base_vec(idx)                                           # [B, n]
        B, n = bits.shape
        out = self.transformer_made(bits)[:, :n, :].reshape(B, n, 2, 2)            # (outcome, re|im)
        rolling = self.masker.compute_rolling_acc_eigs(bits)                         # acc. quantum numbers after 0..n qubits
        words = self.qubit_grouping.cont_mask_words                                  # [n, memo_size] int64 bit words
        memo = pt.stack([self.masker.acc_eigs2memo_idx(rolling[t]) for t in range(n)], dim=1)   # [B, n]
        w = words[pt.arange(n, device=bits.device).view(1, n), memo]
        allowed = pt.stack((w & 1, (w >> 1) & 1), dim=-1).bool()                    # [B, n, 2]
        re = pt.where(allowed, out[..., 0], pt.full_like(out[..., 0], -math.inf))
        re = re - 0.5 * pt.logsumexp(2.0 * re, dim=-1, keepdim=True)
        pick = bits.unsqueeze(-1)
        log_abs = pt.gather(re, -1, pick).squeeze(-1).sum(dim=-1)
        phase = pt.gather(out[..., 1], -1, pick).squeeze(-1).sum(dim=-1)
        return pt.complex(log_abs, phase)

    # ---- reference-style surface -------------------------------------------------------------------------------------------------
    def log_psi_of_indices(self, base_idx: pt.Tensor) -> pt.Tensor:
        if pt.is_grad_enabled() and any(p.requires_grad for p in self._params()):
            _lib.require_cuda(self.device)
            return _TransformerLogPsi.apply(self, base_idx.contiguous().view(-1), *self._params())
        return self.log_psi_kernel(base_idx)

    def log_psi(self, base_vec: pt.Tensor, just_return: bool = False) -> pt.Tensor:
        return self.log_psi_of_indices(self.base_vec2base_idx(base_vec))

    def amplitude(self, base_idx: pt.Tensor) -> pt.Tensor:
        lp = self.log_psi_of_indices(base_idx)
        amps = pt.exp(lp)
        amps.log_psi = lp   # calculations.vmc_loss builds the loss on it (no exp -> log round trip)
        return amps

    def forward(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.amplitude(base_idx)

    @pt.no_grad()
    def cond_log_abs(self, qudit_idx: int = None, base_vec: pt.Tensor = None, return_all_if_made: bool = False,
                     mask: pt.Tensor = None, prefix_idx: pt.Tensor = None) -> pt.Tensor:
        """[B, 2] normalised conditional log|psi| of qubit `qudit_idx` (-inf where masked)."""
        dev = _lib.require_cuda(self.device)
        if prefix_idx is None:
            prefix_idx = self.hilbert_space.base_vec2base_idx(base_vec).view(-1) if base_vec.shape[-1] > 0 else \
                pt.zeros(base_vec.shape[0], dtype=pt.int64, device=dev)
        prefix_idx = prefix_idx.contiguous().view(-1)
        B = prefix_idx.shape[0]
        out = pt.empty((B, 2), dtype=pt.float64, device=dev)
        desc = self._descriptor()
        if self.inference_precision == 'tf32':
            packed = self._packed_weights(desc)
            _lib.check(_lib.lib().anqs_transformer_cond_log_abs_tc(ctypes.byref(desc), _lib.dptr(packed), qudit_idx, _lib.dptr(prefix_idx), B,
                                                                   _lib.dptr(out), _lib.stream_ptr(dev)))
        else:
            _lib.check(_lib.lib().anqs_transformer_cond_log_abs(ctypes.byref(desc), qudit_idx, _lib.dptr(prefix_idx), B, _lib.dptr(out),
                                                                _lib.stream_ptr(dev)))
        return out
