"""Locally decomposable symmetries (reference: nqs/nqs/stochastic/symmetries/*.py).

Same class names, constructor keywords and properties as the reference.  The per-qubit eigenvalue tables
are tiny ((qubit_num, 2) integers), so they are plain host integers here; the masker turns them into the
device tables the kernels read.
"""
import numpy as np
import torch as pt

from .abstract_hilbert_space_object import AbstractHilbertSpaceObject


class AbstractLocallyDecomposableSymmetry(AbstractHilbertSpaceObject):
    """abstract_locally_decomposable_symmetry.py:9-96."""
    is_multiplicative = False
    start_eig = 0
    acc_eig2ordinal_mul_const = 1
    acc_eig2ordinal_add_const = 0
    acc_eig2ordinal_div_const = 1

    def part_eig(self, qubit_idx: int, bit: int) -> int:
        raise NotImplementedError

    def compute_part_eig(self, qubit_idx: int = None, base_vec: pt.Tensor = None) -> pt.Tensor:
        table = pt.tensor([self.part_eig(qubit_idx, 0), self.part_eig(qubit_idx, 1)], dtype=pt.int64, device=base_vec.device)
        return table[base_vec]

    def update_acc_eig(self, qubits_seen: int = None, base_vec: pt.Tensor = None, acc_eig: pt.Tensor = None) -> pt.Tensor:
        part = self.compute_part_eig(qubits_seen, base_vec)
        return acc_eig * part if self.is_multiplicative else acc_eig + part

    def acc_eig2ordinal(self, acc_eig):
        return (acc_eig * self.acc_eig2ordinal_mul_const + self.acc_eig2ordinal_add_const) // self.acc_eig2ordinal_div_const

    def ordinal2acc_eig(self, ordinal):
        return (ordinal * self.acc_eig2ordinal_div_const - self.acc_eig2ordinal_add_const) // self.acc_eig2ordinal_mul_const

    def compute_acc_eig(self, base_vec: pt.Tensor) -> pt.Tensor:
        acc = pt.zeros(base_vec.shape[:-1], dtype=pt.int64, device=base_vec.device) + self.start_eig
        for q in range(base_vec.shape[-1]):
            acc = self.update_acc_eig(q, base_vec[..., q], acc)
        return acc


class AbstractAdditiveSymmetry(AbstractLocallyDecomposableSymmetry):
    is_multiplicative = False
    start_eig = 0


class AbstractMultiplicativeSymmetry(AbstractLocallyDecomposableSymmetry):
    is_multiplicative = True
    start_eig = 1


class ParticleNumberSymmetry(AbstractAdditiveSymmetry):
    """particle_number_symmetry.py:8-60: eigenvalue = number of set bits."""

    def __init__(self, *args, particle_num: int = None, **kwargs):
        super().__init__(*args, **kwargs)
        assert particle_num is not None
        assert 0 <= particle_num <= self.qubit_num
        self.particle_num = particle_num

    spectrum_size = property(lambda self: self.qubit_num + 1)
    ref_eig = property(lambda self: self.particle_num)

    def min_acc_eig(self, qubits_seen: int = None):
        return 0

    def max_acc_eig(self, qubits_seen: int = None):
        return qubits_seen

    def part_eig(self, qubit_idx, bit):
        return bit


class SpinHalfProjectionSymmetry(AbstractAdditiveSymmetry):
    """spin_half_projection_symmetry.py:8-64: +1 per set bit on even positions, -1 on odd positions
    (positions taken through hilbert_space.inv_perm)."""

    def __init__(self, *args, spin: int = None, **kwargs):
        super().__init__(*args, **kwargs)
        assert spin is not None
        assert -self.qubit_num <= spin <= self.qubit_num
        self.spin = spin
        inv_perm = self.inv_perm.cpu().numpy()
        self._signs = [1 if (int(inv_perm[q]) % 2) == 0 else -1 for q in range(self.qubit_num)]
        self.min_acc_eigs = pt.zeros(self.qubit_num + 1, dtype=pt.int64)
        self.max_acc_eigs = pt.zeros(self.qubit_num + 1, dtype=pt.int64)
        for seen in range(1, self.qubit_num + 1):
            up = self._signs[seen - 1] > 0
            self.max_acc_eigs[seen] = self.max_acc_eigs[seen - 1] + (1 if up else 0)
            self.min_acc_eigs[seen] = self.min_acc_eigs[seen - 1] - (0 if up else 1)

    spectrum_size = property(lambda self: (self.qubit_num + 1) // 2 + (self.qubit_num // 2) + 1)
    ref_eig = property(lambda self: self.spin)
    acc_eig2ordinal_add_const = property(lambda self: self.qubit_num // 2)

    def min_acc_eig(self, qubits_seen: int = None):
        return int(self.min_acc_eigs[qubits_seen])

    def max_acc_eig(self, qubits_seen: int = None):
        return int(self.max_acc_eigs[qubits_seen])

    def part_eig(self, qubit_idx, bit):
        return bit * self._signs[qubit_idx]


class Z2Symmetry(AbstractMultiplicativeSymmetry):
    """z2_symmetry.py:8-55: parity of the bits under a Pauli-Z string, eigenvalue in {+1, -1}."""
    acc_eig2ordinal_mul_const = -1
    acc_eig2ordinal_add_const = 1
    acc_eig2ordinal_div_const = 2

    def __init__(self, *args, value: int = None, pauli_z_positions=None, **kwargs):
        super().__init__(*args, **kwargs)
        assert value in (-1, 1, None)
        self.value = value
        self.pauli_z_positions = pauli_z_positions
        mask = np.zeros(self.qubit_num, dtype=np.int64)
        for pos in pauli_z_positions:
            assert 0 <= int(pos) <= self.qubit_num
            mask[int(pos)] = 1
        self.pauli_z_mask = pt.from_numpy(mask).to(self.device)
        self._mask = mask

    spectrum_size = 2
    ref_eig = property(lambda self: self.value)

    def min_acc_eig(self, qubits_seen: int = None):
        return -1

    def max_acc_eig(self, qubits_seen: int = None):
        return 1

    def part_eig(self, qubit_idx, bit):
        return -1 if (self._mask[qubit_idx] and bit) else 1


class IdleSymmetry(AbstractAdditiveSymmetry):
    """idle_symmetry.py: the trivial symmetry (everything is physical)."""
    spectrum_size = 1
    ref_eig = 0

    def min_acc_eig(self, qubits_seen: int = None):
        return 0

    def max_acc_eig(self, qubits_seen: int = None):
        return 0

    def part_eig(self, qubit_idx, bit):
        return 0
