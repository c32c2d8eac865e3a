"""CPU, world_size 2, gloo: bench.py's control flow at world > 1 with the GPU engine replaced by a CPU stand-in.

Round 1 lost its whole 1 -> 8 scaling measurement because rank 0 alone entered dist.py collectives (the secondary
measurements) while the other ranks were already tearing the process group down.  These tests pin the two rules that prevent it:
  * bench.main() issues every collective before destroy_process_group() and runs rank-0-only work after it;
  * the dist.py entry points take world_size=1 ("local mode") so one rank may use them inside a live multi-rank group."""
import io
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from anqs_quantum_chemistry_b200 import dist as adist  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_energy(idx):
    x = idx.view(-1).to(torch.float64)
    return torch.complex(torch.sin(x * 7e-3), 0.1 * torch.cos(x * 3e-3))


class _ToyWF(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.theta = torch.nn.Parameter(torch.randn(4, dtype=torch.float64) * 0.3)

    def amplitude(self, idx):
        phi = torch.cos(idx.view(-1, 1).to(torch.float64) * torch.arange(1, 5, dtype=torch.float64) * 1e-3)
        return torch.exp(torch.complex(phi @ self.theta, torch.zeros(phi.shape[0], dtype=torch.float64)))


class StubEngine:
    """Same methods as bench.GpuEngine; every 'kernel' is a torch CPU expression, every collective is the real dist.py call."""
    backend = 'gloo'

    def __init__(self, args, rank, local_rank, world):
        self.args, self.rank, self.world = args, rank, world
        self.device = torch.device('cpu')

    def setup(self):
        self.n_set = self.args.n_unq * self.world
        lo, hi = adist.shard_bounds(self.n_set, self.world, self.rank)
        self.rows = hi - lo
        self.sizes = [self.args.n_unq] * self.world
        rng = np.random.default_rng(11)
        idx = torch.from_numpy(np.sort(rng.choice(10 ** 6, size=self.n_set, replace=False)).astype(np.int64))
        amps = torch.from_numpy(rng.standard_normal(self.n_set) + 1j * rng.standard_normal(self.n_set))
        self.d_idx, self.d_amps = idx[lo:hi].clone(), amps[lo:hi].clone()
        self.U, self.T, self.kernel_name, self.launches_per_step = 7, 11, 'stub_kernel', 1

    def flush_l2(self):
        pass

    def step(self, clock=None):
        g_idx, g_amps, glo, ghi = adist.all_gather_shards(self.d_idx, self.d_amps, sizes=self.sizes)
        m0 = clock.mark() if clock else None
        eloc = _fake_energy(g_idx[glo:ghi])
        m1 = clock.mark() if clock else None
        mean, var, _ = adist.reduce_energy_stats(adist.local_energy_stats(eloc, g_amps[glo:ghi]))
        return eloc, mean, (m0, m1) if clock else None

    def e2e_step(self):
        self.step()

    e2e_bytes = property(lambda self: (24 * self.rows, 16 * self.rows))

    def conn_per_row(self):
        return 3.0

    def vmc_collective(self):
        # the sharded VMC iteration of the real engine: collectives entered by EVERY rank, inside the live group
        assert dist.is_initialized() and dist.get_world_size() == self.world

        def local_energy(local_idx, local_amps):
            eloc = _fake_energy(local_idx)
            mean, var, norm = adist.reduce_energy_stats(adist.local_energy_stats(eloc, local_amps))
            return eloc, mean, var, norm
        mean, var, loss = adist.ShardedEnergyGradient(_ToyWF(), local_energy)(self.d_idx)
        return {'iters_per_s': 1.0, 'n_gpus': self.world, 'energy_last': float(mean.real)}

    def extras_local(self):
        # round 1's failure mode: rank-0-only code that goes through dist.py.  It must find no live group here ...
        assert not dist.is_initialized(), 'rank-0-only extras must run after destroy_process_group()'

        def local_energy(local_idx, local_amps):
            eloc = _fake_energy(local_idx)
            mean, var, norm = adist.reduce_energy_stats(adist.local_energy_stats(eloc, local_amps), world_size=1)
            return eloc, mean, var, norm
        # ... and uses the explicit local mode anyway
        adist.ShardedEnergyGradient(_ToyWF(), local_energy, world_size=1)(self.d_idx)
        return {'enumeration': {'achieved_gbs': 1.0, 'rows': 1, 'algorithmic_bytes': 'stub', 'kernels': 'stub'}}

    def cpu_baseline(self):
        return {'value': 1.0, 'unit': 'E_loc/s', 'cores': 1, 'kind': 'port', 'sample': 'stub'}


def _bench_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    import bench
    buf = io.StringIO()
    old, sys.stdout = sys.stdout, buf
    try:
        bench.main(['--gpus', str(world), '--steps', '2', '--warmup', '3', '--n-unq', '500', '--extras-at-any-world'], engine_cls=StubEngine)
    finally:
        sys.stdout = old
    assert not dist.is_initialized()
    open(os.path.join(out_dir, f'out{rank}'), 'w').write(buf.getvalue())


@pytest.mark.timeout(180)
def test_bench_main_control_flow_world_size_2(tmp_path):
    """Both ranks return, rank 0 prints exactly one JSON line with the whole-job value, rank 1 prints nothing; the rank-0-only
    extras run after the group is gone (StubEngine.extras_local asserts it)."""
    mp.spawn(_bench_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    out0, out1 = open(tmp_path / 'out0').read().strip(), open(tmp_path / 'out1').read().strip()
    assert out1 == ''
    lines = [l for l in out0.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['n_gpus'] == 2 and line['steps'] == 2 and line['warmup'] == 3
    assert line['config']['sampled_set'] == 1000 and line['scaling'] == 'weak'
    assert line['value'] > 0 and line['e2e']['value'] > 0 and 'secondary' in line
    assert line['secondary']['vmc_iteration_c5_sharded']['n_gpus'] == 2
    assert 'cpu_baseline' not in line  # N = 1 only


def test_bench_main_single_process():
    import bench
    buf = io.StringIO()
    old, sys.stdout = sys.stdout, buf
    env = {k: os.environ.pop(k, None) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    try:
        bench.main(['--gpus', '1', '--steps', '2', '--warmup', '3', '--n-unq', '300'], engine_cls=StubEngine)
    finally:
        sys.stdout = old
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v
    line = json.loads(buf.getvalue().strip())
    assert line['n_gpus'] == 1 and line['cpu_baseline']['kind'] == 'port' and 'roofline_enumeration' in line


def _local_mode_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import datetime
    dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=30))
    try:
        if rank == 0:  # alone inside a live 2-rank group: legal with world_size=1, would block for ever without it
            idx = torch.arange(0, 900, 7, dtype=torch.int64)
            amps = _ToyWF().amplitude(idx).detach()
            g_idx, g_amps, lo, hi = adist.all_gather_shards(idx, amps, world_size=1)
            assert (lo, hi) == (0, idx.shape[0]) and torch.equal(g_idx, idx)
            mean, var, norm = adist.reduce_energy_stats(adist.local_energy_stats(_fake_energy(idx), amps), world_size=1)

            def local_energy(local_idx, local_amps):
                e = _fake_energy(local_idx)
                m, v, n = adist.reduce_energy_stats(adist.local_energy_stats(e, local_amps), world_size=1)
                return e, m, v, n
            m2, _, _ = adist.ShardedEnergyGradient(_ToyWF(), local_energy, world_size=1)(idx)
            assert abs(complex(m2) - complex(mean)) < 1e-14
        dist.barrier()
        open(os.path.join(out_dir, f'local_ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_local_mode_inside_a_live_group(tmp_path):
    mp.spawn(_local_mode_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'local_ok0') and os.path.exists(tmp_path / 'local_ok1')


def test_world_size_override_must_be_one_or_the_group():
    with pytest.raises(AssertionError):
        adist.group_world_size(None, 3)
    assert adist.group_world_size(None, 1) == 1 and adist.group_world_size(None, None) == 1
