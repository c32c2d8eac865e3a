"""anqs_quantum_chemistry_b200 — B200-native VMC inner loop of Exferro/anqs_quantum_chemistry.

Host surface mirrors the reference's `nqs` objects on the hot path; compute runs in
hand-written sm_100a kernels behind the C ABI of libanqs_b200.so (include/anqs_b200.h).
"""
from .constants import BASE_INT_TYPE, BASE_REAL_TYPE, BASE_COMPLEX_TYPE  # noqa: F401
from .hilbert_space import HilbertSpace, SampleTable  # noqa: F401
from .pauli_observable import PauliObservable, PauliArraysOperator, LocalEnergyMetrics  # noqa: F401
from .symmetries import ParticleNumberSymmetry, SpinHalfProjectionSymmetry, Z2Symmetry, IdleSymmetry  # noqa: F401
from .masker import LocallyDecomposableMasker  # noqa: F401
from .qubit_grouping import QubitGrouping, QubitGroupingConfig  # noqa: F401
from .anqs import LogAbsPhaseANQS, ANQSConfig, MLPConfig, LocalSamplingConfig  # noqa: F401
from .calculations import (SamplingConfig, SamplingResult, sample, LocalEnergyCalculationConfig, MonteCarloEstimator,  # noqa: F401
                           LocalEnergyResult, compute_local_energies, vmc_loss, SRConfig, SRMetrics, sr, ProcessGradConfig,
                           process_grad)
from .transformer_anqs import TransformerANQS, TransformerANQSConfig, TransformerMADE  # noqa: F401
