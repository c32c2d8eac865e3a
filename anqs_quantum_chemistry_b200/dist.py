"""Data-parallel sharding of the local-energy pass over the GPUs of one box.

The reference is single-process / single-GPU (energy_opt_exp.py:350); this layer is new design.  Units
(destination samples) are independent given the table {configuration -> amplitude}, and the connection
count per sample is (nearly) constant for a fixed (N_alpha, N_beta), so contiguous equal-row shards balance
the work.  Per pass there is exactly one exchange step:

  1. all_gather of each rank's shard of (packed index int64, amplitude complex128)  -> every rank holds the
     whole sampled set and builds the lookup table locally (24 B per unique sample);
  2. each rank evaluates E_loc for its own rows (fused kernel, no communication);
  3. one all_reduce of the packed statistics [sum w, sum w E, sum w E^2] (5 doubles), w = |psi|^2, from which
     the MonteCarloEstimator mean / variance (compute_local_energies.py:48-62) follow.

One process per GPU, torch.distributed (NCCL on GPUs; gloo works for the host-side logic on CPU).
"""
from typing import Tuple

import torch as pt
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n rows: the first n % world_size ranks get one extra row."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def group_world_size(group=None, world_size: int = None) -> int:
    """Number of ranks a collective of this module spans.  An explicit `world_size` wins over the process group: a caller that
    runs on ONE rank inside an initialised multi-rank job (rank-0-only diagnostics, a single-GPU side computation) passes
    world_size=1 and no collective is issued - a collective entered by one rank alone blocks until the NCCL watchdog aborts."""
    if world_size is not None:
        assert world_size == 1 or (dist.is_initialized() and world_size == dist.get_world_size(group)), \
            'world_size must be 1 (local mode) or the size of the process group'
        return int(world_size)
    return dist.get_world_size(group) if dist.is_initialized() else 1


def all_gather_shards(local_idx: pt.Tensor, local_amps: pt.Tensor, group=None, sizes=None, world_size: int = None):
    """Concatenation over ranks (in rank order) of the per-rank shards.
    Returns (global_idx [N] int64, global_amps [N] complex128, lo, hi) with [lo, hi) this rank's rows.
    `sizes` (shard length of every rank, identical on all ranks) skips the size exchange and its host
    synchronisation; equal shards are gathered straight into the result with no staging copies.
    world_size=1 forces the local (no collective) mode, see group_world_size."""
    world = group_world_size(group, world_size)
    local_idx = local_idx.contiguous().view(-1)
    local_amps = local_amps.contiguous().view(-1)
    if world == 1:
        return local_idx, local_amps, 0, local_idx.shape[0]
    rank = dist.get_rank(group)
    dev = local_idx.device
    if sizes is None:
        sz = pt.zeros(world, dtype=pt.int64, device=dev)
        sz[rank] = local_idx.shape[0]
        dist.all_reduce(sz, group=group)
        sizes = sz.cpu().tolist()
    sizes = [int(v) for v in sizes]
    assert sizes[rank] == local_idx.shape[0]
    lo = sum(sizes[:rank])
    cap = max(sizes)
    if min(sizes) == cap:  # equal shards: two collectives, no packing
        g_idx = pt.empty(world * cap, dtype=pt.int64, device=dev)
        g_amps = pt.empty(world * cap, dtype=pt.complex128, device=dev)
        dist.all_gather_into_tensor(g_idx, local_idx, group=group)
        dist.all_gather_into_tensor(pt.view_as_real(g_amps).view(-1), pt.view_as_real(local_amps).reshape(-1), group=group)
        return g_idx, g_amps, lo, lo + cap
    # ragged shards: one packed buffer per rank, three planes of `cap` doubles [index bits | re | im]
    n_loc = local_idx.shape[0]
    packed = pt.zeros((3, cap), dtype=pt.float64, device=dev)
    packed[0, :n_loc] = local_idx.view(pt.float64)
    packed[1, :n_loc] = local_amps.real
    packed[2, :n_loc] = local_amps.imag
    out = pt.empty((world, 3, cap), dtype=pt.float64, device=dev)
    dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=group)
    parts_idx, parts_amp = [], []
    for r, sz in enumerate(sizes):
        parts_idx.append(out[r, 0, :sz].clone().view(pt.int64))
        parts_amp.append(pt.complex(out[r, 1, :sz], out[r, 2, :sz]))
    return pt.cat(parts_idx), pt.cat(parts_amp), lo, lo + sizes[rank]


def local_energy_stats(eloc: pt.Tensor, amps: pt.Tensor) -> pt.Tensor:
    """Packed partial sums [sum w, Re sum wE, Im sum wE, Re sum wE^2, Im sum wE^2], w = |psi|^2.
    Device tensors: one pass of energy_stats_kernel (anqs_energy_stats; the chain of elementwise products and five reductions
    it replaces cost 0.16 ms per 2^20 rows, 1 % of a bench step).  The expression below is what it computes; it serves the
    host-side tests of this module's collective logic (gloo, CPU tensors) only."""
    if eloc.is_cuda:
        from . import _lib
        dev = _lib.require_cuda(eloc.device)
        assert eloc.dtype == pt.complex128 and amps.dtype == pt.complex128 and eloc.shape == amps.shape and eloc.dim() == 1
        e, a = pt.view_as_real(eloc.contiguous()), pt.view_as_real(amps.contiguous())
        out = pt.empty(5, dtype=pt.float64, device=dev)
        work = _lib._workspace(_lib.lib().anqs_energy_stats_workspace(eloc.shape[0]), dev)
        _lib.check(_lib.lib().anqs_energy_stats(_lib.dptr(e), _lib.dptr(a), eloc.shape[0], _lib.dptr(out), _lib.dptr(work),
                                                _lib.stream_ptr(dev)))
        return out
    w = (amps.real * amps.real + amps.imag * amps.imag)
    we = w * eloc
    wee = we * eloc
    return pt.stack((w.sum(), we.real.sum(), we.imag.sum(), wee.real.sum(), wee.imag.sum()))


def reduce_energy_stats(stats: pt.Tensor, group=None, world_size: int = None):
    """all_reduce of the packed sums, then MonteCarloEstimator semantics with theoretical frequencies
    f = |psi|^2 / sum |psi|^2 (compute_local_energies.py:48-62, 107-113):
    mean = sum f E,  var = sum f (E - mean)^2 = sum f E^2 - mean^2 (complex square, as in the reference)."""
    if group_world_size(group, world_size) > 1:
        dist.all_reduce(stats, group=group)
    norm = stats[0]
    mean = pt.complex(stats[1], stats[2]) / norm
    var = pt.complex(stats[3], stats[4]) / norm - mean * mean
    return mean, var, norm


class ShardedLocalEnergy:
    """Sample-aware local energies of a batch that is sharded over the ranks of `group`."""

    def __init__(self, ham, alpha_num: int, beta_num: int, group=None, sizes=None, world_size: int = None):
        self.sizes = sizes
        self.world_size = world_size
        self.ham = ham
        self.alpha_num = alpha_num
        self.beta_num = beta_num
        self.group = group

    def __call__(self, local_idx: pt.Tensor, local_amps: pt.Tensor):
        """local_idx [n_r] or [n_r,1] int64, local_amps [n_r] complex128: this rank's shard.
        Returns (E_loc of the local rows, mean, var) with mean/var over the global batch."""
        return self.stats(local_idx, local_amps)[:3]

    @pt.no_grad()
    def stats(self, local_idx: pt.Tensor, local_amps: pt.Tensor):
        """As __call__, plus the global normalisation sum |psi|^2: (E_loc local, mean, var, norm)."""
        from .hilbert_space import SampleTable
        g_idx, g_amps, lo, hi = all_gather_shards(local_idx, local_amps, self.group, self.sizes, self.world_size)
        table = SampleTable(g_idx, g_amps)
        eloc, _, _ = self.ham.compute_var_local_energy_proxy(
            unq_batch_as_base_indices=g_idx.view(-1, 1), unq_batch_as_amps=g_amps, coupling_method='ham',
            alpha_num=self.alpha_num, beta_num=self.beta_num, row_start=lo, row_len=hi - lo, table=table)
        mean, var, norm = reduce_energy_stats(local_energy_stats(eloc, g_amps[lo:hi]), self.group, self.world_size)
        return eloc, mean, var, norm


class ShardedEnergyGradient:
    """Energy gradient of a batch that is sharded over the ranks of `group`.

    The loss of the reference (EXP:609), L = 2 Re sum_i f_i log(conj psi_i) (E_i - <E>) with theoretical frequencies
    f_i = |psi_i|^2 / sum_j |psi_j|^2 (CLE:107-113), is a sum over samples once <E> and the normalisation are known; both
    come out of the energy pass already reduced over the ranks.  Every rank back-propagates its own terms through its
    own shard and the flat parameter gradients are summed with ONE all-reduce; parameters are replicated, so all ranks
    take the same optimiser step.  `local_energy` is a callable (local_idx, local_amps) -> (E_loc local, mean, var, norm)
    with global mean / var / norm, e.g. ShardedLocalEnergy(ham, n_alpha, n_beta).stats."""

    def __init__(self, wf, local_energy, group=None, world_size: int = None):
        self.wf, self.local_energy, self.group, self.world_size = wf, local_energy, group, world_size

    def __call__(self, local_idx: pt.Tensor):
        """Leaves the global gradient in p.grad of every parameter of wf; returns (mean, var, loss) over the global batch."""
        wf = self.wf
        params = [p for p in wf.parameters() if p.requires_grad]
        for p in params:
            p.grad = None
        amps = wf.amplitude(local_idx)
        eloc, mean, var, norm = self.local_energy(local_idx, amps.detach())
        with pt.no_grad():
            a = amps.detach()
            seed = (a.real * a.real + a.imag * a.imag) / norm * (eloc - mean)   # f_i (E_i - <E>)
        from .calculations import log_conj_psi
        loss = 2 * (seed * log_conj_psi(amps)).sum().real   # no exp -> log round trip when amps carry their log psi
        loss.backward()
        flat = pt.cat([(p.grad if p.grad is not None else pt.zeros_like(p)).reshape(-1) for p in params] + [loss.detach().reshape(1)])
        if group_world_size(self.group, self.world_size) > 1:
            dist.all_reduce(flat, group=self.group)
        off = 0
        for p in params:
            n = p.numel()
            p.grad = flat[off:off + n].view_as(p).clone()
            off += n
        return mean, var, flat[-1]


@pt.no_grad()
def sharded_sample_stats(wf, sample_num: int, seed: int, world_size: int = None, rank: int = None, group=None,
                         draw_mode: str = 'philox', min_nodes_per_rank: int = 64, gather: bool = True):
    """Count-splitting batch sampling (ANQS:494-525) with the sampling tree sharded by sub-tree.

    Every rank expands the first levels identically (same seed, same draws - they are keyed by the packed prefix of the
    node, not by its position) until the level holds at least min_nodes_per_rank * world_size nodes; rank r then keeps
    the contiguous slice shard_bounds(n_nodes, world_size, r) of that level and expands only its own sub-trees: no
    communication.  The union over ranks is bit-identical to wf.sample_stats(sample_num, seed=seed) on one GPU, in the
    same order (sub-trees of a contiguous slice stay contiguous).  With gather=True the (index, count) shards are
    all-gathered (one collective of 16 B per unique sample) and every rank returns the whole set; otherwise the local shard.
    world_size / rank default to the process group's (pass them explicitly to emulate ranks in one process)."""
    explicit = world_size is not None
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size == 1:  # nothing to shard: the wave function's own call (one host read per call in steady state)
        return wf.sample_stats(sample_num, draw_mode=draw_mode, seed=seed)
    mode = {'rint': 0, 'philox': 1}[draw_mode]
    prefix, counts, memo = wf.sample_stats_root(sample_num)
    sharded = world_size == 1
    for q in range(wf.qubit_grouping.qudit_num):
        if not sharded and prefix.shape[0] >= min_nodes_per_rank * world_size:
            lo, hi = shard_bounds(prefix.shape[0], world_size, rank)
            prefix, counts, memo = prefix[lo:hi].contiguous(), counts[lo:hi].contiguous(), memo[lo:hi].contiguous()
            sharded = True
        prefix, counts, memo = wf.sample_stats_level(q, prefix, counts, memo, mode, seed)
    if not sharded:  # the tree never got wide enough: every rank holds everything, keep a slice
        lo, hi = shard_bounds(prefix.shape[0], world_size, rank)
        prefix, counts = prefix[lo:hi].contiguous(), counts[lo:hi].contiguous()
    if gather and world_size > 1 and dist.is_initialized() and (not explicit or world_size == dist.get_world_size(group)):
        g_idx, g_cnt, _, _ = all_gather_shards(prefix, pt.complex(counts, pt.zeros_like(counts)), group)
        return g_idx.view(-1, 1), g_cnt
    return prefix.view(-1, 1), counts.to(pt.complex128)


def reserve_device_memory(device, nbytes: int = None, headroom: float = 1.6) -> int:
    """Puts ONE block of `nbytes` (default: `headroom` x the peak allocation so far) into torch's caching allocator, so that the
    iterations that follow carve their tensors out of it instead of calling cudaMalloc.  Batch sizes drift from one VMC
    iteration to the next (the number of unique samples; each rank's share of them), and every new maximum is a cudaMalloc;
    with peer access enabled (any NCCL job) one such call was measured at 20 - 160 ms on 4 B200s, against 17 ms for the
    whole iteration.  Call it once after a warm-up iteration; returns the number of bytes reserved."""
    dev = pt.device(device)
    if nbytes is None:
        nbytes = int(headroom * pt.cuda.max_memory_allocated(dev))
    pt.cuda.synchronize(dev)
    pt.cuda.empty_cache()
    block = pt.empty(int(nbytes), dtype=pt.uint8, device=dev)
    del block
    return int(nbytes)
