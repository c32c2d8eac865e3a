"""Dtypes of the hot path (reference: nqs/nqs/base/constants.py:3-15): everything is int64 / float64 / complex128."""
import torch as pt

BASE_INT_TYPE = pt.int64
BASE_REAL_TYPE = pt.double
BASE_COMPLEX_TYPE = pt.cdouble

NEGINF = pt.tensor(-float('inf'), dtype=BASE_REAL_TYPE)
