"""Forwarding base class (reference: nqs/nqs/base/abstract_hilbert_space_object.py:10-83)."""
from typing import Tuple

import torch as pt

from .hilbert_space import HilbertSpace


class AbstractHilbertSpaceObject:
    def __init__(self, *, hilbert_space: HilbertSpace = None):
        super().__init__()
        assert hilbert_space is not None
        self.hilbert_space = hilbert_space

    device = property(lambda self: self.hilbert_space.device)
    qubit_num = property(lambda self: self.hilbert_space.qubit_num)
    rdtype = property(lambda self: self.hilbert_space.rdtype)
    cdtype = property(lambda self: self.hilbert_space.cdtype)
    idx_dtype = property(lambda self: self.hilbert_space.idx_dtype)
    parent_dir = property(lambda self: self.hilbert_space.parent_dir)
    rng_seed = property(lambda self: self.hilbert_space.rng_seed)
    rng = property(lambda self: self.hilbert_space.rng)
    perm_type = property(lambda self: self.hilbert_space.perm_type)
    perm = property(lambda self: self.hilbert_space.perm)
    inv_perm = property(lambda self: self.hilbert_space.inv_perm)

    def base_idx2base_vec(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.hilbert_space.base_idx2base_vec(base_idx=base_idx)

    def base_vec2base_idx(self, base_vec: pt.Tensor) -> pt.Tensor:
        return self.hilbert_space.base_vec2base_idx(base_vec=base_vec)

    def popcount(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.hilbert_space.popcount(base_idx)

    def popcount_(self, base_idx: pt.Tensor) -> pt.Tensor:
        return self.hilbert_space.popcount_(base_idx)

    def sort_base_idx(self, base_idx: pt.Tensor, descending: bool = False):
        return self.hilbert_space.sort_base_idx(base_idx=base_idx, descending=descending)

    def find_a_in_b(self, a: pt.Tensor = None, b: pt.Tensor = None) -> Tuple[pt.Tensor, pt.Tensor]:
        return self.hilbert_space.find_a_in_b(a=a, b=b)
