"""Pieces shared by the autoregressive wave functions (MADE LogAbsPhaseANQS, TransformerANQS).

AutoregressiveSamplerMixin: the two samplers of AbstractANQS - breadth-first count splitting (ANQS:494-662) and Gumbel top-k
(ANQS:664-818) - driven level by level through the k4_sampler.cu kernels.  The host class supplies cond_log_abs(qudit_idx=,
prefix_idx=) -> [B, max_qudit_dim] normalised conditional log|psi|, plus qubit_grouping, masker, max_qudit_dim, device.
ParameterVectorMixin: the flat-gradient helpers SR / process_grad use (ANQS:222-286).
"""
import math
from typing import Tuple

import numpy as np
import torch as pt

from . import _lib
from .constants import BASE_COMPLEX_TYPE


class ParameterVectorMixin:
    @property
    def param_num(self):
        if getattr(self, '_param_num', None) is None:
            self._param_num = sum(p.numel() for p in self.parameters())
        return self._param_num

    @property
    def param_shapes(self):
        return tuple(p.shape for p in self.parameters())

    @property
    def cat_grad_splits(self):
        return tuple(p.numel() for p in self.parameters())

    @property
    def cat_grad(self):
        return pt.cat([(p.grad.data if p.grad is not None else pt.zeros_like(p)).reshape(-1) for p in self.parameters()])

    @cat_grad.setter
    def cat_grad(self, cat_grad: pt.Tensor):
        for p, g in zip(self.parameters(), pt.split(cat_grad, self.cat_grad_splits)):
            p.grad = g.reshape(p.shape)

    def clip_grad_norm(self, value: float = None):
        pt.nn.utils.clip_grad_norm_(self.parameters(), value)



class AutoregressiveSamplerMixin:
    def _init_sampler(self):
        self._next_memo = None
        self.sampler_seed = int(self.rng_seed)
        self._sampler_calls = 0

    def _level_tables(self, q: int):
        qg = self.qubit_grouping
        if self._next_memo is None:
            self._next_memo = [pt.from_numpy(np.ascontiguousarray(t.astype(np.int32))).to(self.device) for t in qg.next_memo_host]
        return qg.cont_mask_words[q], self._next_memo[q]

    def _is_unmasked_level(self, q: int) -> bool:
        """Strategy 'DU' (LocalSamplingConfig(masking_depth > 0), ANQS:45-46): the level is DRAWN from the unmasked conditionals
        (cond_log_abs already returns them for such a qudit) and the unphysical children are thrown away afterwards, with the
        samples they carry (ANQS:605-606 + 653-655; ANQS:708-709 + 804-809).  The count-splitting kernels need nothing extra:
        they draw from whatever conditionals they are given and filter the children by the TRUE continuation mask at the
        end.  The Gumbel level lets every child compete (all-ones mask words) and anqs_sampler_gumbel_select_masked drops the
        unphysical survivors."""
        pattern = getattr(self, 'local_sampling_pattern', None)
        return pattern is not None and pattern[q] == 'DU'

    def _all_ones_mask_words(self):
        if getattr(self, '_ones_words', None) is None:
            self._ones_words = pt.full((self.masker.memo_size,), -1, dtype=pt.int64, device=self.device)
        return self._ones_words

    def _start_memo_idx(self) -> int:
        start = np.array([[sym.start_eig for sym in self.masker.symmetries]], dtype=np.int64)
        return int(self.masker.acc_eigs2memo_idx_np(start)[0])

    def _sampler_seed(self, seed, salt=0):
        if seed is None:
            seed = (self.sampler_seed * 0x9E3779B97F4A7C15 + salt + self._sampler_calls) & 0xFFFFFFFFFFFFFFFF
            self._sampler_calls += 1
        return seed

    @pt.no_grad()
    def sample_stats_level(self, q: int, prefix: pt.Tensor, counts: pt.Tensor, memo: pt.Tensor, mode: int, seed: int):
        """One level of ANQS:593-662 for the live nodes (prefix int64 [B], counts float64 [B], memo int32 [B]): conditional
        probabilities of qudit q, exact multinomial split of every count, ordered compaction of the surviving children.
        The binomial draws are keyed by the node's packed prefix, so a level gives the same children no matter how its
        nodes are spread over calls, ranks or GPUs."""
        dev = _lib.require_cuda(self.device)
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        qg = self.qubit_grouping
        B = prefix.shape[0]
        k, D = qg.qubits_per_qudit[q], qg.qudit_dims_host[q]
        cont_q, next_q = self._level_tables(q)
        cond = self.cond_log_abs(qudit_idx=q, prefix_idx=prefix)
        child = pt.empty((B, D), dtype=pt.float64, device=dev)
        n_child = pt.empty(B, dtype=pt.int64, device=dev)
        single = pt.empty(B, dtype=pt.int8, device=dev)   # single-sample parents hand their child over as one byte
        _lib.check(lib.anqs_sampler_split_level(_lib.dptr(cond), self.max_qudit_dim, k, _lib.dptr(counts), _lib.dptr(memo),
                                                _lib.dptr(cont_q), self.masker.memo_size, B, q, mode, seed, 0, _lib.dptr(prefix),
                                                _lib.dptr(child), _lib.dptr(n_child), _lib.dptr(single), sp))
        offsets = pt.empty(B + 1, dtype=pt.int64, device=dev)
        work = pt.empty(max(1, int(lib.anqs_scan_workspace(B)) // 8), dtype=pt.int64, device=dev)
        _lib.check(lib.anqs_exclusive_scan_i64(_lib.dptr(n_child), _lib.dptr(offsets), B, _lib.dptr(work), sp))
        total = int(offsets[-1].item())
        new_prefix = pt.empty(total, dtype=pt.int64, device=dev)
        new_counts = pt.empty(total, dtype=pt.float64, device=dev)
        new_memo = pt.empty(total, dtype=pt.int32, device=dev)
        if total > 0:
            _lib.check(lib.anqs_sampler_emit_children(_lib.dptr(child), k, qg.qudit_starts[q], _lib.dptr(prefix), _lib.dptr(memo),
                                                      _lib.dptr(cont_q), _lib.dptr(next_q), self.masker.memo_size, B,
                                                      _lib.dptr(offsets), _lib.dptr(single), _lib.dptr(new_prefix), _lib.dptr(new_counts),
                                                      _lib.dptr(new_memo), sp))
        return new_prefix, new_counts, new_memo

    def sample_stats_root(self, sample_num: int):
        dev = _lib.require_cuda(self.device)
        return (pt.zeros(1, dtype=pt.int64, device=dev), pt.tensor([float(sample_num)], dtype=pt.float64, device=dev),
                pt.tensor([self._start_memo_idx()], dtype=pt.int32, device=dev))

    @pt.no_grad()
    def sample_stats(self, sample_num: int, draw_mode: str = 'philox', seed: int = None, exact_levels: bool = False) -> Tuple[pt.Tensor, pt.Tensor]:
        """ANQS:494-525: breadth-first count splitting.  Returns (unique indices [N,1] int64, counts [N] complex128).
        draw_mode 'philox' draws binomials from the counter-based generator (seeded by hilbert_space.rng_seed and a
        per-call counter); 'rint' replaces every draw by its rounded mean (deterministic).
        The first call for a given sample_num reads every level's size back to the host (one synchronisation per qudit); later
        calls size the levels from the previous call's sizes (+25 %) and read the host ONCE, after the last level - the same
        samples either way, because the draws are keyed by the node's prefix and dead rows carry a zero count.  A level that
        outgrows its predicted size makes the call fall back to the exact path (exact_levels=True forces it)."""
        seed = self._sampler_seed(seed)
        mode = {'rint': 0, 'philox': 1}[draw_mode]
        if getattr(self, '_level_hints', None) is None:
            self._level_hints = {}
        hints = self._level_hints.get(sample_num)
        if hints is not None and not exact_levels:
            out = self._sample_stats_predicted(sample_num, mode, seed, hints)
            if out is not None:
                return out
        prefix, counts, memo = self.sample_stats_root(sample_num)
        sizes = []
        for q in range(self.qubit_grouping.qudit_num):
            prefix, counts, memo = self.sample_stats_level(q, prefix, counts, memo, mode, seed)
            sizes.append(int(prefix.shape[0]))
        self._level_hints[sample_num] = sizes
        return prefix.view(-1, 1), counts.to(BASE_COMPLEX_TYPE)

    def _sample_stats_predicted(self, sample_num: int, mode: int, seed: int, hints):
        """All levels with capacities predicted from `hints`, one host read at the end; None when a level overflowed."""
        dev = _lib.require_cuda(self.device)
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        qg = self.qubit_grouping
        Q = qg.qudit_num
        prefix, counts, memo = self.sample_stats_root(sample_num)
        totals = pt.zeros(Q, dtype=pt.int64, device=dev)
        caps = []
        for q in range(Q):
            B = prefix.shape[0]
            k, D = qg.qubits_per_qudit[q], qg.qudit_dims_host[q]
            # every live node carries at least one sample, so a level never holds more than sample_num nodes
            cap = max(1, min(B * D, sample_num, int(hints[q] * 1.25) + 1024))
            caps.append(cap)
            cont_q, next_q = self._level_tables(q)
            cond = self.cond_log_abs(qudit_idx=q, prefix_idx=prefix)
            child = pt.empty((B, D), dtype=pt.float64, device=dev)
            n_child = pt.empty(B, dtype=pt.int64, device=dev)
            single = pt.empty(B, dtype=pt.int8, device=dev)
            _lib.check(lib.anqs_sampler_split_level(_lib.dptr(cond), self.max_qudit_dim, k, _lib.dptr(counts), _lib.dptr(memo),
                                                    _lib.dptr(cont_q), self.masker.memo_size, B, q, mode, seed, 0, _lib.dptr(prefix),
                                                    _lib.dptr(child), _lib.dptr(n_child), _lib.dptr(single), sp))
            offsets = pt.empty(B + 1, dtype=pt.int64, device=dev)
            work = pt.empty(max(1, int(lib.anqs_scan_workspace(B)) // 8), dtype=pt.int64, device=dev)
            _lib.check(lib.anqs_exclusive_scan_i64(_lib.dptr(n_child), _lib.dptr(offsets), B, _lib.dptr(work), sp))
            totals[q:q + 1].copy_(offsets[B:B + 1])
            # rows beyond the level's true size stay dead: count 0 (no children), prefix 0, memo 0
            new_prefix = pt.zeros(cap, dtype=pt.int64, device=dev)
            new_counts = pt.zeros(cap, dtype=pt.float64, device=dev)
            new_memo = pt.zeros(cap, dtype=pt.int32, device=dev)
            _lib.check(lib.anqs_sampler_emit_children_capped(_lib.dptr(child), k, qg.qudit_starts[q], _lib.dptr(prefix), _lib.dptr(memo),
                                                             _lib.dptr(cont_q), _lib.dptr(next_q), self.masker.memo_size, B,
                                                             _lib.dptr(offsets), _lib.dptr(single), cap, _lib.dptr(new_prefix),
                                                             _lib.dptr(new_counts), _lib.dptr(new_memo), sp))
            prefix, counts, memo = new_prefix, new_counts, new_memo
        sizes = totals.cpu().tolist()   # the one host read of the call
        if any(sz > cap for sz, cap in zip(sizes, caps)):
            return None
        self._level_hints[sample_num] = sizes
        n = sizes[-1]
        return prefix[:n].view(-1, 1), counts[:n].to(BASE_COMPLEX_TYPE)

    @pt.no_grad()
    def sample_indices_gumbel(self, sample_num: int, seed: int = None, uniforms=None, compact_levels: bool = None):
        """ANQS:778-818: stochastic-beam (Gumbel top-k) sampling without replacement.  Returns (indices [N,1],
        freqs [N] = model probabilities renormalised over the kept set).  `uniforms`, if given, is a callable
        (level, B, D) -> [B, D] float64 tensor of U(0,1) variates (parity tests).
        compact_levels: drop the masked children after every level like the reference does (sorted top-k and one host read per
        level) instead of carrying them on as dead rows and dropping them once at the end (unsorted top-k - a radix select and an
        ordered compaction, no sort - and one host read per call).  Both give the same SET of samples with the same
        frequencies: the counter-based draws are keyed by the row's packed prefix, not by its position.  The default compacts
        only when `uniforms` are injected, whose shapes follow the reference's compacted, sorted levels; the rows then come
        back in the reference's order (descending Gumbel), otherwise in the order of the last level's candidates."""
        if compact_levels is None:
            compact_levels = uniforms is not None
        dev = _lib.require_cuda(self.device)
        lib, sp = _lib.lib(), _lib.stream_ptr(dev)
        qg = self.qubit_grouping
        seed = self._sampler_seed(seed, salt=0x5bd1e995)
        prefix = pt.zeros(1, dtype=pt.int64, device=dev)
        log_prob = pt.zeros(1, dtype=pt.float64, device=dev)
        gumbel = pt.zeros(1, dtype=pt.float64, device=dev)
        memo = pt.tensor([self._start_memo_idx()], dtype=pt.int32, device=dev)
        for q in range(qg.qudit_num):
            B = prefix.shape[0]
            k, D = qg.qubits_per_qudit[q], qg.qudit_dims_host[q]
            cont_q, next_q = self._level_tables(q)
            cond = self.cond_log_abs(qudit_idx=q, prefix_idx=prefix)
            out_lp = pt.empty((B, D), dtype=pt.float64, device=dev)
            out_g = pt.empty((B, D), dtype=pt.float64, device=dev)
            u = uniforms(q, B, D).to(dev).contiguous() if uniforms is not None else None
            du = self._is_unmasked_level(q)
            _lib.check(lib.anqs_sampler_gumbel_level_keyed(_lib.dptr(cond), self.max_qudit_dim, k, _lib.dptr(log_prob), _lib.dptr(gumbel),
                                                           _lib.dptr(memo), _lib.dptr(self._all_ones_mask_words() if du else cont_q),
                                                           self.masker.memo_size, B, q, seed, 0,
                                                           _lib.dptr(prefix), _lib.dptr(u), _lib.dptr(out_lp), _lib.dptr(out_g), sp))
            flat_g = out_g.view(-1)
            keep = min(sample_num, flat_g.shape[0])
            # ANQS:733: the `keep` largest Gumbels; as the head of the stable descending sort only where the order matters
            top_g, top_i = _lib.topk_f64(flat_g, keep, sorted=compact_levels)
            # ANQS:735-776 in one kernel: children of the kept rows; masked children (gumbel = -inf, ANQS:804-809) sort last
            new_prefix = pt.empty(keep, dtype=pt.int64, device=dev)
            new_memo = pt.empty(keep, dtype=pt.int32, device=dev)
            new_lp = pt.empty(keep, dtype=pt.float64, device=dev)
            new_g = pt.empty(keep, dtype=pt.float64, device=dev)
            n_alive = pt.empty(1, dtype=pt.int32, device=dev)
            _lib.check(lib.anqs_sampler_gumbel_select_masked(_lib.dptr(top_i), _lib.dptr(top_g), keep, k, qg.qudit_starts[q],
                                                             _lib.dptr(prefix), _lib.dptr(memo), _lib.dptr(next_q), _lib.dptr(out_lp),
                                                             _lib.dptr(cont_q if du else None), self.masker.memo_size,
                                                             _lib.dptr(new_prefix), _lib.dptr(new_memo), _lib.dptr(new_lp),
                                                             _lib.dptr(new_g), _lib.dptr(n_alive), sp))
            if compact_levels:
                alive = int(n_alive.item())
                if du and alive < keep:
                    # the dropped rows of an unmasked level sit anywhere among the kept ones: stable partition, alive rows first
                    _, order = _lib.sort_pairs(pt.isinf(new_g).to(pt.int64), None, 0, 1)
                    new_prefix, new_memo, new_lp, new_g = new_prefix[order], new_memo[order], new_lp[order], new_g[order]
                prefix, memo, log_prob, gumbel = new_prefix[:alive], new_memo[:alive], new_lp[:alive], new_g[:alive]
            else:
                # no host read here: masked children are carried on as dead rows (they sort last at every level and mask all
                # of their own children) and dropped once after the last level
                prefix, memo, log_prob, gumbel = new_prefix, new_memo, new_lp, new_g
        if not compact_levels:
            # dead rows (Gumbel = -inf) are scattered among the alive ones: one stable 1-bit radix pass puts the alive rows first
            _, order = _lib.sort_pairs(pt.isinf(gumbel).to(pt.int64), None, 0, 1)
            alive = int(n_alive.item())
            prefix, log_prob = prefix[order[:alive]], log_prob[order[:alive]]
        log_prob = log_prob - pt.logsumexp(log_prob, dim=0)
        return prefix.view(-1, 1), pt.exp(log_prob)

