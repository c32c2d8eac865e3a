"""CPU: the numpy oracle of the wave-function side (oracle/anqs_numpy.py) against the golden vectors produced by the
unmodified reference (tests/golden/anqs_*.npz, oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import anqs_numpy as onp
from oracle.make_golden import made_weights

CASES = ['anqs_n12', 'anqs_n14', 'anqs_n20', 'anqs_n56']


def setup_case(name):
    g = load_golden(name)
    n, ne = int(g['qubit_num']), int(g['particle_num'])
    masks = onp.NumberSpinMasks(n, ne)
    nets = made_weights(n, masks.Q, masks.DM, seed=int(g['weight_seed']))
    return g, masks, nets, onp.masked_weights(nets, masks)


@pytest.mark.parametrize('name', CASES)
def test_mask_tables_match_reference(name):
    g, masks, _, _ = setup_case(name)
    n = masks.n
    memo_ref = np.unpackbits(g['memo'])[:(n + 1) * masks.memo_size].reshape(n + 1, masks.memo_size).astype(bool)
    assert np.array_equal(masks.memo, memo_ref)
    for q in range(masks.Q):
        words = (masks.cont_mask[q].astype(np.uint64) << np.arange(masks.dims[q], dtype=np.uint64)).sum(axis=1, dtype=np.uint64)
        assert np.array_equal(words, g['cont_mask_words'][q])
        assert int((masks.next_memo[q] * masks.cont_mask[q]).sum()) == int(g['next_memo_masked_sum'][q])


@pytest.mark.parametrize('name', CASES)
def test_log_psi_and_cond_match_reference(name):
    g, masks, _, (Wa, ba, Wp, bp) = setup_case(name)
    x = g['samples'].view(np.uint64)
    nphys = int(g['n_phys'])
    lp = onp.log_psi(x, masks, Wa, ba, Wp, bp)
    assert np.abs(lp[:nphys] - g['log_psi'][:nphys]).max() < 1e-12
    amp = np.exp(lp)
    assert np.abs(amp - g['amplitude']).max() < 1e-12
    for key in [k for k in g if k.startswith('cond_log_abs_q')]:
        q = int(key.split('q')[-1])
        c = onp.cond_log_abs(x[:nphys], q, masks, Wa, ba)
        ref = g[key]
        assert np.array_equal(np.isneginf(c), np.isneginf(ref))
        fin = ~np.isneginf(ref)
        assert np.abs(c[fin] - ref[fin]).max() < 1e-12


SAMPLER_CASES = CASES + ['anqs_md1_n20']   # LocalSamplingConfig(masking_depth=1): the last qudit is drawn unmasked


@pytest.mark.parametrize('name', SAMPLER_CASES)
def test_sample_stats_rint_matches_reference(name):
    g, masks, _, (Wa, ba, _, _) = setup_case(name)
    md = int(g['masking_depth']) if 'masking_depth' in g else 0
    idx, cnt = onp.sample_stats_rint(int(g['stats_num']), masks, Wa, ba, masking_depth=md)
    assert np.array_equal(idx.view(np.int64), g['stats_idx'])
    assert np.array_equal(cnt, g['stats_counts'])
    if md == 0:
        assert cnt.sum() == int(g['stats_num'])
    else:
        assert cnt.sum() < int(g['stats_num'])   # samples that fell on unphysical children of the unmasked level are lost (ANQS:653-655)


@pytest.mark.parametrize('name', SAMPLER_CASES)
def test_gumbel_matches_reference(name):
    g, masks, _, (Wa, ba, _, _) = setup_case(name)
    md = int(g['masking_depth']) if 'masking_depth' in g else 0
    urng = np.random.default_rng(int(g['weight_seed']) + 4)
    idx, freqs = onp.sample_gumbel(int(g['gumbel_num']), masks, Wa, ba, lambda q, B, D: urng.random((B, D)), masking_depth=md)
    assert np.array_equal(idx.view(np.int64), g['gumbel_idx'])
    assert np.abs(freqs - g['gumbel_freqs']).max() < 1e-12
