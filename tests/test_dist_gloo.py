"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU shard path (anqs_quantum_chemistry_b200/dist.py):
shard bounds, the all_gather of variable-length (index, amplitude) shards, and the all_reduce of the packed
Monte-Carlo statistics.  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anqs_quantum_chemistry_b200 import dist as adist
from anqs_quantum_chemistry_b200.calculations import MonteCarloEstimator


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        idx = torch.from_numpy(np.sort(rng.integers(0, 2 ** 62, size=n_total, dtype=np.int64)))
        amps = torch.from_numpy(rng.standard_normal(n_total) + 1j * rng.standard_normal(n_total))
        eloc = torch.from_numpy(rng.standard_normal(n_total) + 1j * rng.standard_normal(n_total))
        # deliberately unequal shards: rank 0 gets 1/3, rank 1 the rest
        cut = [0, n_total // 3, n_total][rank:rank + 2] if world == 2 else list(adist.shard_bounds(n_total, world, rank))
        lo, hi = cut
        g_idx, g_amps, glo, ghi = adist.all_gather_shards(idx[lo:hi].clone(), amps[lo:hi].clone())
        assert (glo, ghi) == (lo, hi)
        assert torch.equal(g_idx, idx) and torch.equal(g_amps, amps)
        stats = adist.local_energy_stats(eloc[lo:hi], amps[lo:hi])
        mean, var, norm = adist.reduce_energy_stats(stats)
        w = (amps.conj() * amps)
        est = MonteCarloEstimator(values=eloc, counts=w)
        assert abs(complex(mean) - complex(est.mean)) < 1e-12
        assert abs(complex(var) - complex(est.var)) < 1e-10
        assert abs(float(norm) - float(w.real.sum())) < 1e-9
        # equal shards with the sizes known up front (no size exchange, no staging)
        lo2, hi2 = adist.shard_bounds(1000, world, rank)
        g_idx, g_amps, glo, ghi = adist.all_gather_shards(idx[lo2:hi2].clone(), amps[lo2:hi2].clone(), sizes=[500, 500])
        assert (glo, ghi) == (lo2, hi2) and torch.equal(g_idx, idx[:1000]) and torch.equal(g_amps, amps[:1000])
        # empty shard on one rank
        e_idx = idx[:0] if rank == 0 else idx
        e_amps = amps[:0] if rank == 0 else amps
        g_idx, g_amps, glo, ghi = adist.all_gather_shards(e_idx.clone(), e_amps.clone())
        assert torch.equal(g_idx, idx) and (ghi - glo) == e_idx.shape[0]
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            b = [adist.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_paths():
    idx = torch.arange(10, dtype=torch.int64)
    amps = torch.ones(10, dtype=torch.complex128)
    g_idx, g_amps, lo, hi = adist.all_gather_shards(idx, amps)
    assert torch.equal(g_idx, idx) and (lo, hi) == (0, 10)
    mean, var, norm = adist.reduce_energy_stats(adist.local_energy_stats(torch.full((10,), 2.0 + 1j, dtype=torch.complex128), amps))
    assert abs(complex(mean) - (2 + 1j)) < 1e-14 and abs(complex(var)) < 1e-14 and float(norm) == 10.0


@pytest.mark.timeout(120)
def test_all_gather_and_reduce_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1001, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok0') and os.path.exists(tmp_path / 'ok1')


class _ToyWF(torch.nn.Module):
    """psi(x) = exp(sum_k theta_k phi_k(x) + i sum_k eta_k phi_k(x)) on CPU: enough to exercise the gradient exchange."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.theta = torch.nn.Parameter(torch.randn(5, dtype=torch.float64) * 0.3)
        self.eta = torch.nn.Parameter(torch.randn(5, dtype=torch.float64) * 0.3)

    def amplitude(self, idx):
        x = idx.view(-1, 1).to(torch.float64)
        phi = torch.cos(x * torch.arange(1, 6, dtype=torch.float64) * 1e-3)
        return torch.exp(torch.complex(phi @ self.theta, phi @ self.eta))


def _toy_energy(idx):
    x = idx.view(-1).to(torch.float64)
    return torch.complex(torch.sin(x * 7e-3), 0.1 * torch.cos(x * 3e-3))


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        idx = torch.arange(0, 4000, 13, dtype=torch.int64)
        lo, hi = [0, idx.shape[0] // 3, idx.shape[0]][rank:rank + 2]

        def local_energy(local_idx, local_amps):
            eloc = _toy_energy(local_idx)
            mean, var, norm = adist.reduce_energy_stats(adist.local_energy_stats(eloc, local_amps))
            return eloc, mean, var, norm

        wf = _ToyWF()
        mean, var, loss = adist.ShardedEnergyGradient(wf, local_energy)(idx[lo:hi])
        # single-process reference on the whole batch: EXP:609 with theoretical frequencies
        ref = _ToyWF()
        amps = ref.amplitude(idx)
        est = MonteCarloEstimator(values=_toy_energy(idx), counts=(amps.detach().conj() * amps.detach()))
        from anqs_quantum_chemistry_b200.calculations import vmc_loss
        ref_loss = vmc_loss(amps, est)
        ref_loss.backward()
        assert abs(complex(mean) - complex(est.mean)) < 1e-12
        assert abs(float(loss) - float(ref_loss)) < 1e-12
        for p, q in zip(wf.parameters(), ref.parameters()):
            assert float((p.grad - q.grad).abs().max()) < 1e-12
        open(os.path.join(out_dir, f'grad_ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_energy_gradient_world_size_2(tmp_path):
    """Both ranks end up with the gradient of the reference's loss over the WHOLE batch (one all-reduce of the flat gradient)."""
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'grad_ok0') and os.path.exists(tmp_path / 'grad_ok1')
