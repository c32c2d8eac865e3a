"""CPU: host-side logic of the wave-function drop-ins (tables, parameter layout, initialisation) against the golden
vectors from the reference.  No kernels are launched."""
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_golden
from anqs_quantum_chemistry_b200 import (HilbertSpace, ParticleNumberSymmetry, SpinHalfProjectionSymmetry, Z2Symmetry,
                                         LocallyDecomposableMasker, LogAbsPhaseANQS, ANQSConfig)

CASES = ['anqs_n12', 'anqs_n14', 'anqs_n20', 'anqs_n56', 'anqs_z2_n12', 'anqs_md1_n20']


def build(n, ne, device='cpu', seed=0, z2=(), masking_depth=0):
    from anqs_quantum_chemistry_b200 import Z2Symmetry, LocalSamplingConfig
    tmp = tempfile.mkdtemp(prefix='anqs_test_')
    hs = HilbertSpace(qubit_num=n, device=device, parent_dir=tmp, rng_seed=seed)
    syms = (ParticleNumberSymmetry(hilbert_space=hs, particle_num=ne), SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0))
    syms += tuple(Z2Symmetry(hilbert_space=hs, value=int(v), pauli_z_positions=[i for i in range(n) if (int(m) >> i) & 1]) for v, m in z2)
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=syms)
    torch.manual_seed(seed)
    wf = LogAbsPhaseANQS(hilbert_space=hs, masker=masker,
                         config=ANQSConfig(de_mode='MADE', local_sampling_config=LocalSamplingConfig(masking_depth=masking_depth)))
    return hs, masker, wf


@pytest.mark.parametrize('name', CASES)
def test_tables_and_init_match_reference(name):
    g = load_golden(name)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    z2 = tuple(zip(g['z2_values'].tolist(), g['z2_masks'].tolist()))
    hs, masker, wf = build(n, ne, z2=z2, masking_depth=int(g['masking_depth']))
    memo_ref = np.unpackbits(g['memo'])[:(n + 1) * masker.memo_size].reshape(n + 1, masker.memo_size).astype(bool)
    assert np.array_equal(masker.memo_host, memo_ref)
    qg = wf.qubit_grouping
    assert np.array_equal(qg.cont_mask_words_host, g['cont_mask_words'])
    for q in range(qg.qudit_num):
        assert int((qg.next_memo_host[q] * qg.qudit_idx2cont_mask_mul_table[q].numpy()).sum()) == int(g['next_memo_masked_sum'][q])
    assert wf.param_num == int(g['param_num'])
    # same construction order + same torch seed => the reference's initial weights
    sums = np.array([[float(p.sum()), float((p * p).sum()), float(p.reshape(-1)[0]), float(p.reshape(-1)[-1])] for p in wf.parameters()])
    assert np.allclose(sums, g['init_checksums'], rtol=0, atol=1e-12)
    names = [k for k, _ in wf.named_parameters()]
    assert names[:2] == ['log_abs_subnet.layers.0.weight', 'log_abs_subnet.layers.0.bias']
    assert names[-1] == 'phase_subnet.layers.2.bias'


def test_symmetry_descriptors():
    hs, masker, wf = build(12, 4)
    d = masker.symmetry_descriptors()
    assert d[0].tolist() == [0, 4095, 0, 1, 0, 1, 1, 0]
    assert d[1].tolist() == [0, 0x555, 0xAAA, 1, 6, 1, 13, 0]
    # memo index of a prefix computed from the descriptors equals the reference formula N + (n+1)(Sz + n//2)
    x = 0b101101
    N = bin(x).count('1')
    sz = bin(x & 0x555).count('1') - bin(x & 0xAAA).count('1')
    assert int(masker.acc_eigs2memo_idx_np(np.array([[N, sz]]))[0]) == N + 13 * (sz + 6)


def test_z2_symmetry_tables():
    tmp = tempfile.mkdtemp(prefix='anqs_test_')
    hs = HilbertSpace(qubit_num=8, device='cpu', parent_dir=tmp, rng_seed=0)
    syms = (ParticleNumberSymmetry(hilbert_space=hs, particle_num=4), SpinHalfProjectionSymmetry(hilbert_space=hs, spin=0),
            Z2Symmetry(hilbert_space=hs, value=1, pauli_z_positions=[0, 1, 4, 5]))
    masker = LocallyDecomposableMasker(hilbert_space=hs, symmetries=syms)
    assert masker.memo_size == 9 * 9 * 2
    # brute force: physical = 4 electrons, Sz = 0, even parity on the Z string
    vec = (np.arange(256)[:, None] >> np.arange(8)) & 1
    phys = (vec.sum(1) == 4) & ((vec[:, ::2].sum(1) - vec[:, 1::2].sum(1)) == 0) & ((vec[:, [0, 1, 4, 5]].sum(1) % 2) == 0)
    got = masker.mask(torch.from_numpy(vec)).numpy()
    assert np.array_equal(got, phys)
    d = masker.symmetry_descriptors()
    assert d[2].tolist() == [1, 0b110011, 0, -1, 1, 2, 81, 1]


def test_no_cpu_fallback():
    hs, masker, wf = build(12, 4)
    with pytest.raises(RuntimeError, match='no CPU path'):
        wf.amplitude(torch.zeros((4, 1), dtype=torch.int64))
    with pytest.raises(RuntimeError, match='no CPU path'):
        wf.sample_stats(100)


def test_loss_on_attached_log_psi_equals_the_reference_loss():
    """calculations.vmc_loss with `amps.log_psi` attached (no exp -> log round trip) gives the value and the gradient of the
    reference's 2 Re sum f log(conj psi) (E - <E>) (EXP:609), phases beyond (-pi, pi] included."""
    import torch
    from anqs_quantum_chemistry_b200.calculations import vmc_loss, MonteCarloEstimator
    torch.manual_seed(1)
    n = 200
    theta = torch.randn(n, dtype=torch.float64, requires_grad=True)
    phi = (7.0 * torch.randn(n, dtype=torch.float64)).requires_grad_(True)       # many phases outside the principal branch
    eloc = torch.complex(torch.randn(n, dtype=torch.float64), 0.3 * torch.randn(n, dtype=torch.float64))

    def run(attach):
        lp = torch.complex(theta * 0.5, phi)
        amps = torch.exp(lp)
        if attach:
            amps.log_psi = lp
        est = MonteCarloEstimator(values=eloc, counts=(amps.detach().conj() * amps.detach()))
        loss = vmc_loss(amps, est)
        g = torch.autograd.grad(loss, (theta, phi))
        return loss.detach(), g
    l0, g0 = run(False)
    l1, g1 = run(True)
    assert abs(float(l0) - float(l1)) < 1e-12 * max(1.0, abs(float(l0)))
    for a, b in zip(g0, g1):
        assert float((a - b).abs().max()) < 1e-12


def test_config_keywords_are_the_references_and_nothing_is_swallowed():
    """ANQS:68-109, MLP:13-99, ANQS:20-50: the reference's keyword surface.  What the kernels do not implement raises
    NotImplementedError, unknown keywords raise TypeError (the reference's Config hands them to object.__init__), and the
    defaults are the reference's (de_mode 'NADE', depth 2, width 64, tanh, residuals, bias, masking_depth 0)."""
    import torch.nn as nn
    from anqs_quantum_chemistry_b200.anqs import (ANQSConfig, MLPConfig, WidthConfig, BiasConfig, ActivationConfig, LocalSamplingConfig)
    c = ANQSConfig()
    assert c.de_mode == 'NADE' and c.subtract_mean is True and c.use_sign_structure is False
    assert c.local_sampling_config.masking_depth == 0 and c.local_sampling_config.strategy == 'MU'
    m = c.main_subnet_config
    assert (m.depth, m.width, m.use_res, m.use_bias, m.activation, m.activate_last_layer) == (2, 64, True, True, nn.Tanh, False)
    m = MLPConfig(depth=3, width_config=WidthConfig(width=64), bias_config=BiasConfig(use_bias=False),
                  activation_config=ActivationConfig(activation=nn.Tanh), use_res=False)
    assert (m.depth, m.width, m.use_bias, m.use_res) == (3, 64, False, False)
    assert m.width_config.create_pattern(3) == (64, 64, 64)
    for bad in (dict(width_config=WidthConfig(width=128)), dict(activation_config=ActivationConfig(activation=nn.ReLU)),
                dict(activate_last_layer=True), dict(depth=5), dict(width=32)):
        with pytest.raises(NotImplementedError):
            MLPConfig(**bad)
    with pytest.raises(NotImplementedError):
        WidthConfig(pattern_type='custom')
    with pytest.raises(NotImplementedError):
        ANQSConfig(use_sign_structure=True)
    for cls, kw in ((MLPConfig, dict(hidden=64)), (ANQSConfig, dict(mode='MADE')), (WidthConfig, dict(widths=(64,))),
                    (LocalSamplingConfig, dict(depth=1))):
        with pytest.raises(TypeError):
            cls(**kw)
    with pytest.raises(TypeError):
        MLPConfig(width=64, width_config=WidthConfig())
    assert LocalSamplingConfig(masking_depth=2).create_local_sampling_pattern(qudit_num=4) == ('MU', 'MU', 'DU', 'DU')
    with pytest.raises(AssertionError):
        LocalSamplingConfig(masking_depth=5).create_local_sampling_pattern(qudit_num=4)
