"""QubitGrouping drop-in (reference: nqs/nqs/base/qubit_grouping.py:13-213).

Groups qubits into qudits of k bits (default 6) and tabulates, per qudit and per accumulated-quantum-number
index (`memo_idx`), the memo index after each of the D = 2^k local outcomes and whether that continuation can
still reach the physical sector (QG:99-108).  Tables are built on the host with numpy (Q x memo_size x D
entries) and uploaded; the kernels read the continuation masks as one 64-bit word per (qudit, memo_idx).
"""
from typing import Tuple

import numpy as np
import torch as pt

from .abstract_hilbert_space_object import AbstractHilbertSpaceObject
from .hilbert_space import HilbertSpace
from .masker import LocallyDecomposableMasker


class QubitGroupingConfig:
    FIELDS = ('type', 'qubit_per_qudit')

    def __init__(self, *args, type: str = 'uniform', qubit_per_qudit: int = 6, **kwargs):
        self.type = type
        self.qubit_per_qudit = qubit_per_qudit


class QubitGrouping(AbstractHilbertSpaceObject):
    def __init__(self, *args, qudit_starts: Tuple[int] = None, qudit_ends: Tuple[int] = None,
                 masker: LocallyDecomposableMasker = None, **kwargs):
        super().__init__(*args, **kwargs)
        assert len(qudit_starts) == len(qudit_ends)
        for q in range(len(qudit_starts)):
            assert 0 <= qudit_starts[q] <= self.qubit_num - 1
            assert 1 <= qudit_ends[q] <= self.qubit_num
            assert qudit_starts[q] < qudit_ends[q]
        self.qudit_num = len(qudit_starts)
        self.qudit_starts = tuple(int(v) for v in qudit_starts)
        self.qudit_ends = tuple(int(v) for v in qudit_ends)
        self.qubits_per_qudit = tuple(e - s for s, e in zip(self.qudit_starts, self.qudit_ends))
        assert max(self.qubits_per_qudit) <= 6, 'the kernels hold one continuation mask per 64-bit word: qubit_per_qudit <= 6'
        dims = tuple(2 ** k for k in self.qubits_per_qudit)
        self.qudit_dims_host = dims
        dev = self.device
        self.qudit_dims = pt.tensor(dims, device=dev)

        two_power, q_of_qubit = [], []
        for q in range(self.qudit_num):
            two_power += [2 ** j for j in range(self.qubits_per_qudit[q])]
            q_of_qubit += [q] * self.qubits_per_qudit[q]
        self.qubit_idx2qudit_two_power = pt.tensor(two_power, dtype=self.idx_dtype, device=dev)
        self.qubit_idx2qudit_idx = pt.tensor(q_of_qubit, dtype=self.idx_dtype, device=dev)

        self.masker = masker
        local_vecs, local_eigs = [], []
        for q in range(self.qudit_num):  # QG:75-96
            k, D = self.qubits_per_qudit[q], dims[q]
            vec = (np.arange(D, dtype=np.int64).reshape(-1, 1) >> np.arange(k, dtype=np.int64)) & 1
            acc = np.broadcast_to(masker.host['start'], (D, masker.sym_num)).copy()
            for j in range(k):
                acc = masker.update_acc_eigs_np(self.qudit_starts[q] + j, vec[:, j], acc)
            local_vecs.append(vec)
            local_eigs.append(acc)
        self.qudit_idx2local_base_vecs = tuple(pt.from_numpy(v).to(dev) for v in local_vecs)
        self.qudit_idx2local_eigs = tuple(pt.from_numpy(e).to(dev) for e in local_eigs)

        # QG:98-108: multiplication tables over every memo index
        M = masker.memo_size
        self.memo_idx_arange = pt.arange(M, device=dev)
        all_eigs = masker.memo_idx2acc_eigs_np(np.arange(M))
        self.memo_idx_acc_eigs = pt.from_numpy(all_eigs).to(dev)
        mult = masker.host['is_multiplicative']
        next_tables, mask_tables = [], []
        mask_words = np.zeros((self.qudit_num, M), np.uint64)
        for q in range(self.qudit_num):
            D = dims[q]
            new = np.where(mult, all_eigs[:, None, :] * local_eigs[q][None, :, :], all_eigs[:, None, :] + local_eigs[q][None, :, :])
            inb = masker.bound_check_np(self.qudit_ends[q], new)
            idx = masker.acc_eigs2memo_idx_np(new)
            mask = np.zeros((M, D), bool)
            mask[inb] = masker.memo_host[self.qudit_ends[q], idx[inb]]
            next_tables.append(idx)
            mask_tables.append(mask)
            mask_words[q] = (mask.astype(np.uint64) << np.arange(D, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
        self.qudit_idx2memo_idx_mul_table = [pt.from_numpy(t).to(dev) for t in next_tables]
        self.qudit_idx2cont_mask_mul_table = [pt.from_numpy(t).to(dev) for t in mask_tables]
        self.cont_mask_words_host = mask_words                        # [Q, memo_size] uint64, bit d = outcome d allowed
        self.next_memo_host = next_tables                             # Q x [memo_size, D] int64
        self._cont_mask_words = None

    @property
    def cont_mask_words(self) -> pt.Tensor:
        """[Q, memo_size] int64 on the device (bit d of word = continuation d allowed)."""
        if self._cont_mask_words is None:
            self._cont_mask_words = pt.from_numpy(self.cont_mask_words_host.view(np.int64)).to(self.device)
        return self._cont_mask_words

    @classmethod
    def create(cls, config: QubitGroupingConfig = None, hs: HilbertSpace = None, masker: LocallyDecomposableMasker = None):
        config = config if config is not None else QubitGroupingConfig()
        if config.type == 'uniform':
            k = config.qubit_per_qudit
            qudit_num = hs.qubit_num // k + (1 if hs.qubit_num % k else 0)
            starts = tuple(q * k for q in range(qudit_num))
            ends = starts[1:] + (hs.qubit_num,)
            return QubitGrouping(hilbert_space=hs, qudit_starts=starts, qudit_ends=ends, masker=masker)
        raise RuntimeError(f'Wrong qubit grouping type: {config.type}')

    @staticmethod
    def qudit2base_vec(qudit: pt.Tensor, qubit_per_qudit: int = None) -> pt.Tensor:
        shifts = pt.arange(0, qubit_per_qudit, dtype=qudit.dtype, device=qudit.device)
        return (qudit.reshape(-1, 1) >> shifts).remainder_(2)

    @staticmethod
    def base_vec2qudit(base_vec: pt.Tensor, qubit_per_qudit: int = None) -> pt.Tensor:
        assert base_vec.shape[-1] == qubit_per_qudit
        return pt.sum(base_vec * (2 ** pt.arange(qubit_per_qudit, dtype=base_vec.dtype, device=base_vec.device)), dim=-1)

    def base_vec2qudit_base_vec(self, base_vec: pt.Tensor) -> pt.Tensor:
        return pt.scatter_add(pt.zeros((base_vec.shape[0], self.qudit_num), dtype=self.idx_dtype, device=base_vec.device),
                              dim=1, index=pt.broadcast_to(self.qubit_idx2qudit_idx, base_vec.shape),
                              src=base_vec * self.qubit_idx2qudit_two_power)

    def base_vec2qudit_rolling_acc_eigs(self, base_vec: pt.Tensor) -> Tuple[pt.Tensor]:
        rolling = self.masker.compute_rolling_acc_eigs(base_vec)
        return (rolling[0],) + tuple(rolling[self.qudit_ends[q]] for q in range(self.qudit_num))

    def base_vec2qudit_continuation_masks(self, base_vec: pt.Tensor = None) -> Tuple[pt.Tensor]:
        rolling = self.base_vec2qudit_rolling_acc_eigs(base_vec)
        memo_idx = [self.masker.acc_eigs2memo_idx(e) for e in rolling]
        return tuple(pt.reshape(self.qudit_idx2cont_mask_mul_table[q][memo_idx[q]], (-1,)) for q in range(self.qudit_num))
