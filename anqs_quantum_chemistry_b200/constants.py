"""Dtypes of the hot path (reference: nqs/nqs/base/constants.py:3-15): everything is int64 / float64 / complex128.
Only the names callers of the reference import are kept; its 32-bit index variant is not supported here."""
import torch

BASE_INT_TYPE, BASE_REAL_TYPE, BASE_COMPLEX_TYPE = torch.int64, torch.float64, torch.complex128
NEGINF = torch.full((), float('-inf'), dtype=BASE_REAL_TYPE)   # 0-dim, like the reference's constant
