"""Call sites of the hot path (reference: nqs/nqs/applications/quantum_chemistry/experiments/calculations/
{sample,compute_local_energies}.py).  Same names, keywords and return shapes, so a driver written against the
reference (EnergyOptExp.iter, energy_opt_exp.py:626-679) runs unchanged on the B200 objects.

  sample(wf, config)                    SMP:51-101   -> SamplingResult, unique count, repetitions, next sample_num
  compute_local_energies(wf, ...)       CLE:75-163   -> LocalEnergyResult(full / sample-aware MonteCarloEstimator), metrics
  MonteCarloEstimator                   CLE:48-62    mean / var under frequencies counts / sum(counts)
  vmc_loss(...)                         EXP:609      2 Re sum f log(conj psi) (E - <E>)
  sr(wf, sampling_result, config)       SR:88-136    stochastic reconfiguration of wf.cat_grad on the most frequent samples
  process_grad(wf, ...)                 PG:55-70     SR, gradient clipping, renormalisation
"""
import numpy as np
import torch as pt

from . import _lib


class SamplingConfig:
    """SMP:18-36."""
    FIELDS = ('sample_indices', 'sample_num', 'sample_precisely', 'upscale_factor', 'downscale_factor', 'couple_spin_flip')

    def __init__(self, *args, sample_indices: bool = True, sample_num: int = 10000, sample_precisely: bool = False,
                 upscale_factor: float = 3.0, downscale_factor: float = 2.0, couple_spin_flip: bool = False, **kwargs):
        self.sample_indices, self.sample_num, self.sample_precisely = sample_indices, sample_num, sample_precisely
        self.upscale_factor, self.downscale_factor, self.couple_spin_flip = upscale_factor, downscale_factor, couple_spin_flip


class SamplingResult:
    """SMP:39-48."""

    def __init__(self, *args, indices: pt.Tensor = None, counts: pt.Tensor = None, **kwargs):
        self.indices, self.counts = indices, counts


def sample(wf=None, config: SamplingConfig = None, starting_sample_num: int = None, verbose: bool = False, **sampler_kwargs):
    """SMP:51-101.  `sampler_kwargs` (seed=, draw_mode=, uniforms=) are forwarded to the wave function's sampler."""
    repetition_num = 1
    next_rep_sample_num = starting_sample_num
    if config.sample_indices:
        indices, counts = wf.sample_indices_gumbel(sample_num=config.sample_num, **sampler_kwargs)
        next_rep_sample_num = config.sample_num
        actual_unq_num = indices.shape[0]
    elif config.sample_precisely:
        # grow the number of samples until at least sample_num unique configurations come back, keep the top ones
        while True:
            indices, counts = wf.sample_stats(sample_num=int(next_rep_sample_num), **sampler_kwargs)
            if indices.shape[0] > config.sample_num:
                if next_rep_sample_num / config.downscale_factor > config.sample_num:
                    next_rep_sample_num /= config.downscale_factor
                break
            if indices.shape[0] == config.sample_num:
                break
            repetition_num += 1
            next_rep_sample_num *= config.upscale_factor
        actual_unq_num = indices.shape[0]
        counts, order = _lib.sort_pairs(counts.real.contiguous(), None, 0, 64, key_kind=1, xor_mask=-1)  # descending, stable
        indices = indices[order[:config.sample_num]]
    else:
        indices, counts = wf.sample_stats(sample_num=config.sample_num, **sampler_kwargs)
        next_rep_sample_num = config.sample_num
        actual_unq_num = indices.shape[0]
    counts = counts.type(wf.cdtype).to(wf.device)
    if verbose:
        print(f'sampled {indices.shape[0]} unique configurations carrying {counts.sum()} samples')
    if config.couple_spin_flip:
        with pt.no_grad():
            indices = pt.cat((indices, wf.spin_flip_base_idx(indices)), dim=0)
            indices, _ = wf.hilbert_space.compute_unique_indices(indices)
            counts = wf.amplitude(indices)
            counts = pt.conj(counts) * counts
            counts = counts / pt.sum(counts)
    return SamplingResult(indices=indices, counts=counts), actual_unq_num, repetition_num, next_rep_sample_num


class LocalEnergyCalculationConfig:
    """CLE:25-45."""
    ALLOWED_CODE_VERSIONS = ('old', 'new')

    def __init__(self, *args, use_theor_freqs: bool = True, use_tree_for_candidates: str = 'all_to_all',
                 sampled_indices_chunk_size: int = 20000, amps_chunk_size: int = 100000, code_version: str = 'new',
                 matrix_element_chunk_size=np.inf, **kwargs):
        assert code_version in self.ALLOWED_CODE_VERSIONS
        self.use_theor_freqs, self.use_tree_for_candidates = use_theor_freqs, use_tree_for_candidates
        self.sampled_indices_chunk_size, self.amps_chunk_size = sampled_indices_chunk_size, amps_chunk_size
        self.code_version, self.matrix_element_chunk_size = code_version, matrix_element_chunk_size


class MonteCarloEstimator:
    """CLE:48-62: freqs = counts / sum(counts); mean = values . freqs; var = (values - mean)^2 . freqs (complex square)."""

    def __init__(self, *args, values: pt.Tensor = None, counts: pt.Tensor = None, **kwargs):
        self.values = values
        self.freqs = counts / pt.sum(counts)
        if self.values is not None and self.freqs is not None:
            self.mean = pt.dot(self.values, self.freqs)
            self.var = pt.dot(pt.pow(self.values - self.mean, 2), self.freqs)
        else:
            self.mean = self.var = None


class LocalEnergyResult:
    def __init__(self, *args, full_e_loc_mc_est: MonteCarloEstimator = None, sample_aware_e_loc_mc_est: MonteCarloEstimator = None,
                 **kwargs):
        self.full_e_loc_mc_est, self.sample_aware_e_loc_mc_est = full_e_loc_mc_est, sample_aware_e_loc_mc_est


@pt.no_grad()
def compute_local_energies(wf=None, sampling_result: SamplingResult = None, sampled_amps: pt.Tensor = None, ham=None,
                           config: LocalEnergyCalculationConfig = None, sample_aware: bool = True, verbose: bool = False):
    """CLE:75-163.  The reference's code_version='new' has no full (not sample-aware) energy (CLE:93-94) and
    its 'old' version is the only route to it; here both versions accept both modes and give the same numbers."""
    config = config if config is not None else LocalEnergyCalculationConfig()
    if config.use_tree_for_candidates not in ('ham', 'all_to_all', 'trie'):
        raise ValueError(f'Wrong coupling mode: {config.use_tree_for_candidates}')
    indices = sampling_result.indices
    theor_freqs = pt.conj(sampled_amps) * sampled_amps
    theor_freqs = theor_freqs / pt.sum(theor_freqs)
    if sample_aware:
        n_el = wf.masker.symmetries[0].particle_num
        full, aware, metrics = ham.compute_var_local_energy_proxy(
            unq_batch_as_base_indices=indices, unq_batch_as_amps=sampled_amps, coupling_method=config.use_tree_for_candidates,
            chunk_size=config.sampled_indices_chunk_size, alpha_num=n_el // 2, beta_num=n_el // 2,
            matrix_element_chunk_size=config.matrix_element_chunk_size)
    else:
        full, aware, metrics = ham.compute_local_energies(wf=wf, sampled_indices=indices, sampled_amps=sampled_amps, verbose=verbose,
                                                          sample_aware=False, chunk_size=config.sampled_indices_chunk_size,
                                                          compute_via_ham_xy_coupling=True, amps_chunk_size=config.amps_chunk_size)
    full_counts = theor_freqs if config.use_theor_freqs else sampling_result.counts
    return LocalEnergyResult(full_e_loc_mc_est=MonteCarloEstimator(values=full, counts=full_counts),
                             sample_aware_e_loc_mc_est=MonteCarloEstimator(values=aware, counts=theor_freqs)), metrics


def log_conj_psi(sampled_amps: pt.Tensor) -> pt.Tensor:
    """log(conj psi_i) of the loss EXP:609.  The reference exponentiates log psi in `amplitude`
    (ANQS:485) and takes the logarithm again in the loss, so the backward pass runs through exp and log (a division by psi)
    for nothing.  The wave functions of this package attach the log psi they computed to the amplitudes they return
    (`amps.log_psi`, same autograd graph); when it is there the loss is built on it directly: conj(log psi), with the
    imaginary part brought back to the principal branch by a constant (no gradient) multiple of 2 pi so that the VALUE of the
    loss equals the reference's log(conj(psi)) too.  Amplitudes from anywhere else take the reference's route."""
    lp = getattr(sampled_amps, 'log_psi', None)
    if lp is None or lp.shape != sampled_amps.shape:
        return pt.log(pt.conj(sampled_amps))
    a = -lp.imag
    with pt.no_grad():
        wrap = -2.0 * np.pi * pt.round(a / (2.0 * np.pi))
    return pt.complex(lp.real, a + wrap)


def vmc_loss(sampled_amps: pt.Tensor, estimator: MonteCarloEstimator) -> pt.Tensor:
    """EXP:609: 2 Re sum_i f_i log(conj psi_i) (E_i - <E>); its gradient is the energy gradient."""
    return 2 * (estimator.freqs * log_conj_psi(sampled_amps) * (estimator.values - estimator.mean)).sum().real


# ---- gradient post-processing ------------------------------------------------------------------
class SRConfig:
    """SR:20-32."""

    def __init__(self, *args, max_indices_num: int = 25, use_theor_freqs: bool = False, use_reg: bool = True, reg_eps: float = 1e-4,
                 **kwargs):
        self.max_indices_num, self.use_theor_freqs, self.use_reg, self.reg_eps = max_indices_num, use_theor_freqs, use_reg, reg_eps


class SRMetrics:
    FIELDS = ('sr_unq_num', 'sr_sampled_prob', 'sr_max_amp', 'sr_min_amp', 'sr_time')

    def __init__(self, **kwargs):
        for f in self.FIELDS:
            setattr(self, f, kwargs.get(f, np.nan))


@pt.no_grad()
def sr(wf=None, sampling_result: SamplingResult = None, config: SRConfig = None):
    """SR:88-136.  Preconditions the gradient g with the quantum geometric tensor S = O^dagger O of the max_indices_num most
    frequent samples, O = diag(sqrt f) conj(J - <J>), J the per-sample log-Jacobian (wf.compute_cat_log_jac):
      regularised:  g <- (g - O^dagger (1 + eps T)^-1 (O g)) / eps  with T = O O^dagger / eps^2  (Woodbury form of (S + eps)^-1 g)
      otherwise:    g <- O^dagger T^+ T^+ O g  (pseudo-inverse through an SVD).
    Returns (new real gradient, SRMetrics, seconds) like the reference's @timed function."""
    import time
    t0 = time.time()
    config = config if config is not None else SRConfig()
    metrics = SRMetrics()
    g = wf.cat_grad
    g = pt.complex(g, pt.zeros_like(g))
    if config.use_theor_freqs:
        amps = wf.amplitude(sampling_result.indices)
        freqs = pt.conj(amps) * amps
    else:
        freqs = sampling_result.counts
    freqs = freqs / pt.sum(freqs)
    # SR:96-101 sorts all the frequencies and keeps the first max_indices_num: the radix select finds them without the full sort
    _, top = _lib.topk_f64(freqs.real.contiguous(), min(int(config.max_indices_num), freqs.shape[0]), sorted=True)
    f = freqs[top]
    f = f / pt.sum(f)
    idx = sampling_result.indices[top]
    amps = wf.amplitude(idx)
    metrics.sr_unq_num = amps.shape[0]
    # the three metrics of SR:104-107 in one device-to-host copy instead of three
    host = pt.stack((pt.dot(pt.conj(amps), amps), amps[0], amps[-1])).cpu()
    metrics.sr_sampled_prob = host[0].real.item()
    metrics.sr_max_amp, metrics.sr_min_amp = host[1].item(), host[2].item()
    J = wf.compute_cat_log_jac(idx)
    J = J - (J.T * f).sum(dim=-1, keepdim=True).T
    if config.use_reg:
        inv_eps = 1.0 / config.reg_eps
        O = inv_eps * pt.sqrt(f).unsqueeze(1) * J.conj()
        T = O @ O.conj().T
        T_reg = pt.eye(T.shape[0], dtype=T.dtype, device=T.device) + config.reg_eps * T
        g = inv_eps * g - O.conj().T @ pt.linalg.solve(T_reg, O @ g)
    else:
        O = pt.sqrt(f).unsqueeze(1) * J.conj()
        T = O @ O.conj().T
        u, sv, vh = pt.linalg.svd(T, full_matrices=True)
        sv_inv = pt.where(pt.isclose(sv, pt.zeros_like(sv)), pt.zeros_like(sv), 1.0 / sv)
        T_inv = vh.conj().T @ pt.diag(sv_inv.type(T.dtype)) @ u.conj().T
        g = O.conj().T @ (T_inv.conj().T @ (T_inv @ (O @ g)))
    return g.real, metrics, time.time() - t0


class ProcessGradConfig:
    """PG:20-34."""

    def __init__(self, *args, use_sr: bool = True, sr_config: SRConfig = None, clip_grad_norm: bool = True,
                 clip_grad_norm_value: float = 1.0, renorm_grad: bool = False, **kwargs):
        self.use_sr = use_sr
        self.sr_config = sr_config if sr_config is not None else SRConfig()
        self.clip_grad_norm, self.clip_grad_norm_value, self.renorm_grad = clip_grad_norm, clip_grad_norm_value, renorm_grad


class ProcessGradMetrics:
    def __init__(self):
        self.sr_metrics, self.proc_grad_time = SRMetrics(), np.nan


def process_grad(wf=None, sampling_result: SamplingResult = None, config: ProcessGradConfig = None):
    """PG:55-70."""
    config = config if config is not None else ProcessGradConfig()
    metrics = ProcessGradMetrics()
    if config.use_sr:
        wf.cat_grad, metrics.sr_metrics, sr_time = sr(wf=wf, sampling_result=sampling_result, config=config.sr_config)
        metrics.sr_metrics.sr_time = sr_time
    if config.clip_grad_norm:
        wf.clip_grad_norm(config.clip_grad_norm_value)
    if config.renorm_grad:
        wf.cat_grad = wf.cat_grad / pt.linalg.norm(wf.cat_grad)
    return metrics
