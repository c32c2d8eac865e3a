"""LocallyDecomposableMasker drop-in (reference: nqs/nqs/stochastic/maskers/locally_decomposable_masker.py:17-177).

The boolean DP table `memo[qubits_seen, memo_idx]` ("can a prefix with these accumulated quantum numbers still
reach the reference sector?") has (n+1) x memo_size entries (57 x 3249 at 56 qubits), so it is built on the
host with numpy once, cached under parent_dir/memos like the reference (MSK:62-65), and uploaded.  The
kernels never walk it: they read the per-qudit continuation bitmasks QubitGrouping derives from it.
"""
import os
from typing import Tuple

import numpy as np
import torch as pt

from .abstract_hilbert_space_object import AbstractHilbertSpaceObject
from .symmetries import AbstractLocallyDecomposableSymmetry


class LocallyDecomposableMasker(AbstractHilbertSpaceObject):
    def __init__(self, *args, symmetries: Tuple[AbstractLocallyDecomposableSymmetry] = None, **kwargs):
        super().__init__(*args, **kwargs)
        for sym in symmetries:
            assert isinstance(sym, AbstractLocallyDecomposableSymmetry)
        self.symmetries = tuple(symmetries)
        self.sym_num = len(self.symmetries)
        n, S = self.qubit_num, self.sym_num

        self.memo_size = 1
        h = dict(is_multiplicative=np.zeros(S, bool), bases=np.zeros(S, np.int64), ref_acc_eigs=np.zeros(S, np.int64),
                 mul=np.zeros(S, np.int64), add=np.zeros(S, np.int64), div=np.zeros(S, np.int64),
                 start=np.zeros(S, np.int64), local_eigs=np.zeros((n, 2, S), np.int64),
                 min_bounds=np.zeros((n + 1, S), np.int64), max_bounds=np.zeros((n + 1, S), np.int64))
        for s, sym in enumerate(self.symmetries):  # MSK:42-61
            h['is_multiplicative'][s] = sym.is_multiplicative
            h['bases'][s] = self.memo_size
            self.memo_size *= int(sym.spectrum_size)
            h['ref_acc_eigs'][s] = sym.ref_eig
            h['mul'][s], h['add'][s], h['div'][s] = (sym.acc_eig2ordinal_mul_const, sym.acc_eig2ordinal_add_const,
                                                     sym.acc_eig2ordinal_div_const)
            h['start'][s] = sym.start_eig
            for seen in range(n + 1):
                h['min_bounds'][seen, s] = sym.min_acc_eig(seen)
                h['max_bounds'][seen, s] = sym.max_acc_eig(seen)
            for q in range(n):
                h['local_eigs'][q, 0, s] = sym.part_eig(q, 0)
                h['local_eigs'][q, 1, s] = sym.part_eig(q, 1)
        self.host = h
        dev = self.device
        self.is_multiplicative = pt.from_numpy(h['is_multiplicative']).to(dev)
        self.bases = pt.from_numpy(h['bases']).to(dev)
        self.ref_acc_eigs = pt.from_numpy(h['ref_acc_eigs']).to(dev)
        self.acc_eig2ordinal_mul_consts = pt.from_numpy(h['mul']).to(dev)
        self.acc_eig2ordinal_add_consts = pt.from_numpy(h['add']).to(dev)
        self.acc_eig2ordinal_div_consts = pt.from_numpy(h['div']).to(dev)
        self.local_eigs = pt.from_numpy(h['local_eigs']).to(dev)
        self.min_bounds = pt.from_numpy(h['min_bounds']).to(dev)
        self.max_bounds = pt.from_numpy(h['max_bounds']).to(dev)

        memos_dir = os.path.join(self.parent_dir, 'memos')
        os.makedirs(memos_dir, exist_ok=True)
        self.memo_filename = os.path.join(memos_dir, f'{self.perm_type}_{(n + 1, self.memo_size)}_{h["ref_acc_eigs"].tolist()}.npy')
        self.init_memo()

    # ---- host (numpy) arithmetic: MSK:67-108 ---------------------------------------------------------
    def acc_eigs2memo_idx_np(self, acc_eigs: np.ndarray) -> np.ndarray:
        h = self.host
        return (((acc_eigs * h['mul'] + h['add']) // h['div']) * h['bases']).sum(axis=-1)

    def memo_idx2acc_eigs_np(self, memo_idx: np.ndarray) -> np.ndarray:
        h = self.host
        memo_idx = np.array(memo_idx, dtype=np.int64, copy=True)
        out = np.zeros(memo_idx.shape + (self.sym_num,), np.int64)
        for s in range(self.sym_num - 1, -1, -1):
            ordinal = memo_idx // h['bases'][s]
            memo_idx -= h['bases'][s] * ordinal
            out[..., s] = (ordinal * h['div'][s] - h['add'][s]) // h['mul'][s]
        return out

    def update_acc_eigs_np(self, qubit_idx: int, bit, acc_eigs: np.ndarray) -> np.ndarray:
        h = self.host
        local = h['local_eigs'][qubit_idx][bit]
        return np.where(h['is_multiplicative'], acc_eigs * local, acc_eigs + local)

    def bound_check_np(self, qubits_seen: int, acc_eigs: np.ndarray) -> np.ndarray:
        h = self.host
        return np.all((acc_eigs >= h['min_bounds'][qubits_seen]) & (acc_eigs <= h['max_bounds'][qubits_seen]), axis=-1)

    def init_memo(self):
        """Backward DP of MSK:130-146."""
        if os.path.exists(self.memo_filename):
            memo = np.load(self.memo_filename)
        else:
            n = self.qubit_num
            memo = np.zeros((n + 1, self.memo_size), bool)
            eigs = self.memo_idx2acc_eigs_np(np.arange(self.memo_size))
            memo[n] = (eigs == self.host['ref_acc_eigs']).all(axis=-1)
            for seen in range(n - 1, -1, -1):
                ok = np.zeros(self.memo_size, bool)
                for bit in (0, 1):
                    nxt = self.update_acc_eigs_np(seen, bit, eigs)
                    inb = self.bound_check_np(seen + 1, nxt)
                    idx = self.acc_eigs2memo_idx_np(nxt)
                    phys = np.zeros(self.memo_size, bool)
                    phys[inb] = memo[seen + 1, idx[inb]]
                    ok |= phys
                memo[seen] = self.bound_check_np(seen, eigs) & ok
            np.save(self.memo_filename, memo)
        self.memo_host = memo
        self.memo = pt.from_numpy(memo).to(self.device)

    # ---- reference tensor surface ------------------------------------------------------------------------
    def acc_eigs2memo_idx(self, acc_eigs: pt.Tensor = None) -> pt.Tensor:
        if not pt.is_tensor(acc_eigs):
            acc_eigs = pt.tensor(acc_eigs, dtype=self.idx_dtype, device=self.device)
        memo_idx = (acc_eigs * self.acc_eig2ordinal_mul_consts + self.acc_eig2ordinal_add_consts) // self.acc_eig2ordinal_div_consts
        return (memo_idx * self.bases).sum(dim=-1)

    def memo_idx2acc_eigs(self, memo_idx: pt.Tensor = None) -> pt.Tensor:
        if not pt.is_tensor(memo_idx):
            memo_idx = pt.tensor(memo_idx, dtype=self.idx_dtype, device=self.device)
        return pt.from_numpy(self.memo_idx2acc_eigs_np(memo_idx.cpu().numpy())).to(memo_idx.device)

    def update_acc_eigs(self, qubit_idx: int = None, base_vec: pt.Tensor = None, acc_eigs: pt.Tensor = None) -> pt.Tensor:
        local = self.local_eigs[qubit_idx, base_vec, :]
        return pt.where(pt.unsqueeze(self.is_multiplicative, dim=0), acc_eigs * local, acc_eigs + local)

    def acc_eigs_bound_check(self, qubits_seen: int = None, acc_eigs: pt.Tensor = None):
        return pt.all(pt.logical_and(pt.greater_equal(acc_eigs, self.min_bounds[qubits_seen, :]),
                                     pt.less_equal(acc_eigs, self.max_bounds[qubits_seen, :])), dim=-1)

    def compute_rolling_acc_eigs(self, base_vec: pt.Tensor = None) -> Tuple[pt.Tensor]:
        start = pt.from_numpy(self.host['start']).to(base_vec.device)
        rolling = [start.expand(*base_vec.shape[:-1], self.sym_num).clone()]
        for seen in range(base_vec.shape[-1]):
            rolling.append(self.update_acc_eigs(seen, base_vec[..., seen], rolling[seen]))
        return tuple(rolling)

    def mask(self, base_vec: pt.Tensor) -> pt.Tensor:
        acc = pt.stack([sym.compute_acc_eig(base_vec) for sym in self.symmetries], dim=-1)
        return self.memo[base_vec.shape[-1], self.acc_eigs2memo_idx(acc)]

    # ---- what the kernels consume --------------------------------------------------------------------------
    def symmetry_descriptors(self) -> np.ndarray:
        """One row of 8 int64 per symmetry: (kind, plus_mask, minus_mask, mul, add, div, base, start).
        kind 0 (additive): eig(prefix) = start + popcount(prefix & plus) - popcount(prefix & minus);
        kind 1 (multiplicative, Z2): eig(prefix) = start * (-1)^popcount(prefix & plus).
        Bit i of the packed configuration is base_vec[:, i] (HS:130)."""
        h = self.host
        rows = np.zeros((self.sym_num, 8), np.int64)
        for s in range(self.sym_num):
            plus = minus = 0
            for q in range(self.qubit_num):
                e0, e1 = int(h['local_eigs'][q, 0, s]), int(h['local_eigs'][q, 1, s])
                if h['is_multiplicative'][s]:
                    assert e0 == 1 and e1 in (1, -1), 'multiplicative symmetries must have local eigenvalues in {+1, -1}'
                    plus |= (1 << q) if e1 == -1 else 0
                else:
                    assert e0 == 0 and e1 in (-1, 0, 1), 'additive symmetries must have local eigenvalues in {-1, 0, +1}'
                    plus |= (1 << q) if e1 == 1 else 0
                    minus |= (1 << q) if e1 == -1 else 0
            to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
            rows[s] = (int(h['is_multiplicative'][s]), to_i64(plus), to_i64(minus), h['mul'][s], h['add'][s], h['div'][s],
                       h['bases'][s], h['start'][s])
        return rows
