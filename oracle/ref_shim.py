"""TEST INFRASTRUCTURE ONLY — import shim for running the *unmodified* reference on CPU.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import
anything under `oracle/`.  This module additionally needs `/root/reference`, which exists
only in the build container (never on the GPU box): it is used by
`oracle/make_golden.py` to generate the fixtures committed under `tests/golden/` and by the
container-only tests marked `needs_reference`.

The reference cannot be imported as-is:
  * nqs/nqs/utils/custom_popcount/cuda_int64popcount.py:8-9 evaluates
    `pt.cuda.current_stream()` in a class body at import time (needs a CUDA driver);
  * cuda_int64popcount.py:28,67 call `cupy.RawKernel` at module level (cupy not installed);
  * nqs/nqs/stochastic/observables/pauli_observable.py:9-10 import openfermion.
The shim stubs exactly those three things and nothing else; the reference code that runs is
the reference's own, with `popcount_mode='memory_efficient'` so the CuPy stub is never called.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get('ANQS_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'nqs', 'nqs'))


class QubitOperator:
    """Stand-in for openfermion.QubitOperator: the reference only reads `.terms`
    (pauli_observable.py:101,157): {((qubit, 'X'|'Y'|'Z'), ...): coefficient}, () = identity."""

    def __init__(self, terms=None):
        self.terms = dict(terms or {})


def count_qubits(op) -> int:
    return 1 + max((q for t in op.terms for q, _ in t), default=-1)


_LOADED = {}


def load_reference():
    """Returns a namespace with the reference classes used by the hot path."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')

    if not torch.cuda.is_available():
        class _Stream:
            cuda_stream = 0
        torch.cuda.current_stream = lambda *a, **k: _Stream()

    if 'cupy' not in sys.modules:
        cupy = types.ModuleType('cupy')

        class _RawKernel:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                raise RuntimeError("CuPy kernel stub called: use popcount_mode='memory_efficient'")
        cupy.RawKernel = _RawKernel
        sys.modules['cupy'] = cupy

    if 'openfermion' not in sys.modules:
        of = types.ModuleType('openfermion')
        ofu = types.ModuleType('openfermion.utils')
        ofu.count_qubits = count_qubits
        of.QubitOperator = QubitOperator
        of.utils = ofu
        sys.modules['openfermion'] = of
        sys.modules['openfermion.utils'] = ofu

    path = os.path.join(REFERENCE_ROOT, 'nqs')
    if path not in sys.path:
        sys.path.insert(0, path)

    import nqs.base  # noqa: F401  must precede nqs.utils (circular import in the reference)
    from nqs.base import HilbertSpace
    from nqs.base.qubit_grouping import QubitGrouping, QubitGroupingConfig
    from nqs.stochastic.observables.pauli_observable import PauliObservable
    from nqs.stochastic.symmetries import ParticleNumberSymmetry, SpinHalfProjectionSymmetry
    from nqs.stochastic.symmetries.z2_symmetry import Z2Symmetry
    from nqs.stochastic.maskers import LocallyDecomposableMasker
    from nqs.stochastic.ansatzes.anqs import ANQSConfig, LogAbsPhaseANQS, LogPsiANQS
    from nqs.stochastic.ansatzes.anqs.abstract_anqs import AbstractANQS, LocalSamplingConfig
    from nqs.stochastic.ansatzes.anqs.mlp import MLP, MLPConfig
    from nqs.applications.quantum_chemistry.experiments.calculations import (
        SamplingConfig, SamplingResult, sample, LocalEnergyCalculationConfig,
        compute_local_energies, ProcessGradConfig, SRConfig, process_grad)

    _LOADED.update(dict(
        HilbertSpace=HilbertSpace, QubitGrouping=QubitGrouping, QubitGroupingConfig=QubitGroupingConfig,
        PauliObservable=PauliObservable, ParticleNumberSymmetry=ParticleNumberSymmetry,
        SpinHalfProjectionSymmetry=SpinHalfProjectionSymmetry, Z2Symmetry=Z2Symmetry,
        LocallyDecomposableMasker=LocallyDecomposableMasker, ANQSConfig=ANQSConfig,
        LogAbsPhaseANQS=LogAbsPhaseANQS, LogPsiANQS=LogPsiANQS, AbstractANQS=AbstractANQS,
        LocalSamplingConfig=LocalSamplingConfig, MLP=MLP, MLPConfig=MLPConfig,
        SamplingConfig=SamplingConfig, SamplingResult=SamplingResult, sample=sample,
        LocalEnergyCalculationConfig=LocalEnergyCalculationConfig,
        compute_local_energies=compute_local_energies, ProcessGradConfig=ProcessGradConfig,
        SRConfig=SRConfig, process_grad=process_grad, QubitOperator=QubitOperator))
    return types.SimpleNamespace(**_LOADED)


def build_reference_objects(terms: dict, qubit_num: int, particle_num: int, parent_dir: str,
                            de_mode: str = 'MADE', rng_seed: int = 0, spin: int = 0, z2=(), masking_depth: int = 0):
    """Wires HilbertSpace -> PauliObservable -> masker -> LogAbsPhaseANQS exactly like
    energy_opt_exp.py:348-376 with the 'e_num_spin' masker level (create_masker.py:61-65)."""
    ref = load_reference()
    os.makedirs(parent_dir, exist_ok=True)
    hs = ref.HilbertSpace(qubit_num=qubit_num, device=torch.device('cpu'), parent_dir=parent_dir,
                          rng_seed=rng_seed, popcount_mode='memory_efficient')
    hs.init_perm()
    ham = ref.PauliObservable(hilbert_space=hs, of_qubit_operator=ref.QubitOperator(terms)) if terms is not None else None
    # z2: ((value, pauli_z_positions), ...) -> Z2Symmetry generators on top of the two additive symmetries (the reference's
    # default masker level 'z2', create_masker.py:18-24, 58-69); masking_depth: LocalSamplingConfig (ANQS:29-50)
    syms = (ref.ParticleNumberSymmetry(hilbert_space=hs, particle_num=particle_num),
            ref.SpinHalfProjectionSymmetry(hilbert_space=hs, spin=spin))
    syms += tuple(ref.Z2Symmetry(hilbert_space=hs, value=v, pauli_z_positions=torch.tensor(list(pos), dtype=torch.int64)) for v, pos in z2)
    masker = ref.LocallyDecomposableMasker(hilbert_space=hs, symmetries=syms)
    torch.manual_seed(rng_seed)
    wf = ref.LogAbsPhaseANQS(hilbert_space=hs, masker=masker, config=ref.ANQSConfig(
        de_mode=de_mode, local_sampling_config=ref.LocalSamplingConfig(masking_depth=masking_depth)))
    return types.SimpleNamespace(ref=ref, hs=hs, ham=ham, masker=masker, wf=wf)
