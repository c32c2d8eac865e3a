"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's MADE wave function, symmetry masks and the two
autoregressive samplers (float64 / int64, CPU).  The product never imports this module.

Parity status: pinned against the unmodified reference through tests/golden/anqs_*.npz (made by
oracle/make_golden.py via oracle/ref_shim.py): masks and tables bit-exact, log psi / conditional log-amplitudes /
gradients to 1e-12, sample_stats (with the binomial draw replaced by its rounded mean on both sides) and the
Gumbel top-k sampler (same uniforms on both sides) bit-exact in the sampled configurations.

Every function cites the reference lines it restates (paths relative to /root/reference/nqs/nqs/):
  ANQS = stochastic/ansatzes/anqs/abstract_anqs.py   LAP = stochastic/ansatzes/anqs/log_abs_phase_anqs.py
  MLP  = stochastic/ansatzes/anqs/mlp.py             QG  = base/qubit_grouping.py
  MSK  = stochastic/maskers/locally_decomposable_masker.py   SYM = stochastic/symmetries/
"""
import numpy as np


# ---- symmetry masks: particle number + S_z (SYM/particle_number_symmetry.py:8-60, spin_half_projection_symmetry.py:8-64)
class NumberSpinMasks:
    """memo[(qubits_seen, memo_idx)] DP of MSK:130-146 and the per-qudit tables of QG:99-108 for the
    'e_num_spin' masker (create_masker.py:61-65): memo_idx = N + (n+1) * (S_z + n//2) (MSK:67-73)."""

    def __init__(self, n, particle_num, spin=0, qubit_per_qudit=6):
        self.n, self.particle_num, self.spin = n, particle_num, spin
        self.base = n + 1
        self.memo_size = (n + 1) * ((n + 1) // 2 + n // 2 + 1)
        sign = np.array([1 if q % 2 == 0 else -1 for q in range(n)])          # spin_half_projection_symmetry.py:52
        self.sign = sign
        max_sz = np.concatenate(([0], np.cumsum(sign > 0)))                      # :17-28
        min_sz = -np.concatenate(([0], np.cumsum(sign < 0)))
        self.min_sz, self.max_sz = min_sz, max_sz
        N = np.arange(self.memo_size) % self.base
        Sz = np.arange(self.memo_size) // self.base - n // 2
        memo = np.zeros((n + 1, self.memo_size), bool)
        memo[n] = (N == particle_num) & (Sz == spin)                              # MSK:136
        for seen in range(n - 1, -1, -1):                                         # MSK:137-145
            ok = np.zeros(self.memo_size, bool)
            for bit in (0, 1):
                N2, Sz2 = N + bit, Sz + bit * sign[seen]
                inb = (N2 >= 0) & (N2 <= seen + 1) & (Sz2 >= min_sz[seen + 1]) & (Sz2 <= max_sz[seen + 1])
                idx = N2 + self.base * (Sz2 + n // 2)
                phys = np.zeros(self.memo_size, bool)
                phys[inb] = memo[seen + 1, idx[inb]]
                ok |= phys
            memo[seen] = (N >= 0) & (N <= seen) & (Sz >= min_sz[seen]) & (Sz <= max_sz[seen]) & ok
        self.memo = memo
        # qudits (QG:111-128)
        k = qubit_per_qudit
        Q = n // k + (1 if n % k else 0)
        self.starts = [q * k for q in range(Q)]
        self.ends = self.starts[1:] + [n]
        self.bits = [e - s for s, e in zip(self.starts, self.ends)]
        self.dims = [2 ** b for b in self.bits]
        self.Q, self.DM = Q, max(self.dims)
        self.cont_mask, self.next_memo = [], []
        for q in range(Q):                                                       # QG:98-108, 167-197
            D = self.dims[q]
            d = np.arange(D)
            dN = np.zeros(D, np.int64)
            dS = np.zeros(D, np.int64)
            for j in range(self.bits[q]):
                b = (d >> j) & 1
                dN += b
                dS += b * sign[self.starts[q] + j]
            N2, Sz2 = N[:, None] + dN[None, :], Sz[:, None] + dS[None, :]
            end = self.ends[q]
            inb = (N2 >= 0) & (N2 <= end) & (Sz2 >= min_sz[end]) & (Sz2 <= max_sz[end])
            idx = N2 + self.base * (Sz2 + n // 2)
            mask = np.zeros((self.memo_size, D), bool)
            mask[inb] = memo[end, idx[inb]]
            self.cont_mask.append(mask)
            self.next_memo.append(idx)

    def memo_idx_of_prefix(self, x, length):
        """MSK:156-167 + 67-73 for the first `length` bits of packed configurations x (uint64 array)."""
        x = np.asarray(x, dtype=np.uint64)
        N = np.zeros(x.shape, np.int64)
        Sz = np.zeros(x.shape, np.int64)
        for q in range(length):
            b = ((x >> np.uint64(q)) & np.uint64(1)).astype(np.int64)
            N += b
            Sz += b * self.sign[q]
        return N + self.base * (Sz + self.n // 2)


# ---- the masked MLP (MLP:217-246) ------------------------------------------------------------------------------
def mlp_forward(x, Ws, bs, use_res=True):
    """tanh hidden layers with residual adds on layers 1..depth-1 (MLP:237-239), identity output (MLP:144-148).
    Ws are the weights AFTER multiplication by the MADE masks (MLP:230-233)."""
    depth = len(Ws) - 1
    hs = []
    for l in range(depth + 1):
        y = x @ Ws[l].T + (bs[l] if bs[l] is not None else 0.0)
        if use_res and 0 < l < depth:
            y = y + x
        x = np.tanh(y) if l < depth else y
        if l < depth:
            hs.append(x)
    return x, hs


def encode(x, n, known=None):
    """1 - 2*bit for the known positions, 0 beyond the prefix (MLP:205-225)."""
    x = np.asarray(x, dtype=np.uint64)
    bits = ((x[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.float64)
    v = 1.0 - 2.0 * bits
    if known is not None:
        v[:, known:] = 0.0
    return v


def _normalise(z, mask):
    """where(mask, z, -inf) then z - 0.5*logsumexp(2z) with nan -> -inf (ANQS:358-364, 392-405)."""
    zm = np.where(mask, z, -np.inf)
    with np.errstate(invalid='ignore', divide='ignore'):
        mx = zm.max(axis=-1, keepdims=True)
        lse = mx * 2 + np.log(np.exp(2 * zm - 2 * mx).sum(axis=-1, keepdims=True))
        out = zm - 0.5 * lse
    return np.where(np.isnan(out), -np.inf, out)


def cond_log_abs(x_prefix, q, masks: NumberSpinMasks, W_abs, b_abs, use_res=True, subtract_mean=True, du=False):
    """LAP:105-163 (MADE branch) for qudit q: [B, DM]."""
    n, DM = masks.n, masks.DM
    x_prefix = np.asarray(x_prefix, dtype=np.uint64)
    y, _ = mlp_forward(encode(x_prefix, n, known=masks.starts[q]), W_abs, b_abs, use_res)
    y = y.reshape(-1, masks.Q, DM)[:, q, :]
    if subtract_mean:
        y = y - y.mean(axis=-1, keepdims=True)                                   # before masking (ANQS:338-340)
    mi = masks.memo_idx_of_prefix(x_prefix, masks.starts[q])
    m = np.zeros((x_prefix.shape[0], DM), bool)
    m[:, :masks.dims[q]] = masks.cont_mask[q][mi]
    if du == 'sampler':
        m[:, :masks.dims[q]] = True                                              # the samplers' all-ones mask has the qudit's own width (ANQS:605-613)
    elif du:
        m[:] = True                                                              # log_psi unmasks max_qudit_dim columns (ANQS:417-418, 434-441)
    return _normalise(y, m)


def log_psi(x, masks: NumberSpinMasks, W_abs, b_abs, W_ph, b_ph, use_res=True, subtract_mean=True):
    """ANQS:407-481 (MADE branch) with LAP:63-103: complex128 [B]."""
    n, Q, DM = masks.n, masks.Q, masks.DM
    x = np.asarray(x, dtype=np.uint64)
    inp = encode(x, n)
    ya, _ = mlp_forward(inp, W_abs, b_abs, use_res)
    yp, _ = mlp_forward(inp, W_ph, b_ph, use_res)
    ya, yp = ya.reshape(-1, Q, DM), yp.reshape(-1, Q, DM)
    if subtract_mean:
        ya = ya - ya.mean(axis=-1, keepdims=True)
    re = np.zeros(x.shape[0])
    im = np.zeros(x.shape[0])
    rows = np.arange(x.shape[0])
    for q in range(Q):
        mi = masks.memo_idx_of_prefix(x, masks.starts[q])
        m = np.zeros((x.shape[0], DM), bool)
        m[:, :masks.dims[q]] = masks.cont_mask[q][mi]
        c = ((x >> np.uint64(masks.starts[q])) & np.uint64(masks.dims[q] - 1)).astype(np.int64)
        cond = _normalise(ya[:, q, :], m)
        re = re + cond[rows, c]
        im = im + np.where(m[rows, c], np.pi * yp[rows, q, c], 0.0)
    im = np.where(np.isneginf(re), 0.0, im)
    return re + 1j * im


# ---- count-splitting sampler (ANQS:494-525, 557-662) ---------------------------------------------------------------
def split_counts_rint(cond, counts, k):
    """sample_mult_new_new (ANQS:557-591) with Binomial(n, p).sample() replaced by rint(n p): [B, 2^k] counts."""
    D = 1 << k
    logits = 2.0 * cond[:, :D]
    with np.errstate(invalid='ignore', divide='ignore'):
        mx = logits.max(axis=-1, keepdims=True)
        e = np.exp(logits - mx)
        p = e / e.sum(axis=-1, keepdims=True)
    p = np.nan_to_num(p, nan=0.0)
    cum = np.concatenate((np.zeros((p.shape[0], 1)), np.cumsum(p, axis=-1)), axis=-1)
    out = np.zeros((cond.shape[0], D))
    node = {0: np.asarray(counts, dtype=np.float64)}
    for j in range(k):
        span = D >> j
        nxt = {}
        for t, cnt in node.items():
            lo, mid, hi = t * span, t * span + span // 2, (t + 1) * span
            succ, fail = cum[:, mid] - cum[:, lo], cum[:, hi] - cum[:, mid]
            with np.errstate(invalid='ignore', divide='ignore'):
                pr = np.nan_to_num(succ / (succ + fail), nan=0.0)
            left = np.minimum(cnt, np.maximum(0.0, np.rint(cnt * pr)))
            nxt[2 * t], nxt[2 * t + 1] = left, cnt - left
        node = nxt
    for d, cnt in node.items():
        out[:, d] = cnt
    return out


def sample_stats_rint(sample_num, masks: NumberSpinMasks, W_abs, b_abs, use_res=True, subtract_mean=True, masking_depth=0):
    """ANQS:494-525 with deterministic draws: (indices uint64 [N], counts float64 [N]) in the reference's order.
    masking_depth: the last `masking_depth` qudits are drawn from the UNMASKED conditionals (strategy 'DU', ANQS:45-46, 605-606)
    and their unphysical children dropped afterwards with the samples they carry (ANQS:653-655)."""
    n = masks.n
    prefix = np.zeros(1, np.uint64)
    counts = np.array([float(sample_num)])
    memo = np.array([0 + masks.base * (0 + n // 2)])
    for q in range(masks.Q):
        D = masks.dims[q]
        cond = cond_log_abs(prefix, q, masks, W_abs, b_abs, use_res, subtract_mean, du='sampler' if q >= masks.Q - masking_depth else False)
        child = split_counts_rint(cond, counts, masks.bits[q])
        allowed = masks.cont_mask[q][memo]
        keep = allowed & (child > 0)
        b, d = np.nonzero(keep)                                                # (parent, outcome) order (ANQS:645-660)
        prefix = prefix[b] | (d.astype(np.uint64) << np.uint64(masks.starts[q]))
        counts = child[b, d]
        memo = masks.next_memo[q][memo[b], d]
    return prefix, counts


# ---- Gumbel top-k sampler (ANQS:664-818) ----------------------------------------------------------------------------
def _log1mexp(x):
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.where(x > -0.693, np.log(-np.expm1(x)), np.log1p(-np.exp(x)))


def _log1pexp(x):
    with np.errstate(invalid='ignore', over='ignore'):
        return np.where(x < 18.0, np.log1p(np.exp(x)), x + np.exp(-x))


def sample_gumbel(sample_num, masks: NumberSpinMasks, W_abs, b_abs, uniform_fn, use_res=True, subtract_mean=True, masking_depth=0):
    """ANQS:778-818; uniform_fn(level, B, D) supplies what pt.rand((B, D)) returned in the reference run.  masking_depth as in
    sample_stats_rint (ANQS:708-709: unphysical children of an unmasked level compete in the top-k and are dropped after it)."""
    n = masks.n
    prefix = np.zeros(1, np.uint64)
    lp = np.zeros(1)
    G = np.zeros(1)
    memo = np.array([0 + masks.base * (0 + n // 2)])
    for q in range(masks.Q):
        D = masks.dims[q]
        cond = cond_log_abs(prefix, q, masks, W_abs, b_abs, use_res, subtract_mean, du='sampler' if q >= masks.Q - masking_depth else False)[:, :D]
        with np.errstate(invalid='ignore', divide='ignore'):
            phi = np.nan_to_num(lp[:, None] + 2.0 * cond, nan=-np.inf, neginf=-np.inf, posinf=np.inf)
            u = uniform_fn(q, prefix.shape[0], D)
            g = phi - np.log(-np.log(u))
            Z = g.max(axis=-1, keepdims=True)
            v = G[:, None] - g + _log1mexp(g - Z)
            out = G[:, None] - np.maximum(v, 0.0) - _log1pexp(-np.abs(v))
        out = np.where(np.isnan(out), -np.inf, out).reshape(-1)
        order = np.argsort(-out, kind='stable')[:sample_num]
        b, d = order // D, order % D
        new_memo = masks.next_memo[q][memo[b], d]
        phys = masks.cont_mask[q][memo[b], d]                                   # == memo[end, new_memo] (ANQS:804-809)
        b, d, order, new_memo = b[phys], d[phys], order[phys], new_memo[phys]
        prefix = prefix[b] | (d.astype(np.uint64) << np.uint64(masks.starts[q]))
        lp, G, memo = phi.reshape(-1)[order], out[order], new_memo
    lp = lp - (lp.max() + np.log(np.exp(lp - lp.max()).sum()))
    return prefix, np.exp(lp)


# ---- MADE causal masks (MLP:170-203) ---------------------------------------------------------------------------------
def made_masks(ends, Q, DM, depth=2, width=64):
    """[start_mask (width x n), mid masks (width x width) ..., end_mask (Q*DM x width)] as float64 0/1 arrays."""
    allowed = []
    for _l in range(depth):
        row = []
        for g in range(Q):
            row += [g] * (width // Q + 1 * ((Q - g - 1) < (width % Q)))
        allowed.append(np.array(row))
    start_connect = []
    for g in range(Q):
        start_connect += [g] * (ends[g] - (ends[g - 1] if g > 0 else 0))
    start_connect = np.array(start_connect)
    masks = [(allowed[0][:, None] > start_connect[None, :]).astype(np.float64)]
    for l in range(1, depth):
        masks.append((allowed[l][:, None] >= allowed[l - 1][None, :]).astype(np.float64))
    end_connect = np.repeat(np.arange(Q), DM)
    masks.append((end_connect[:, None] >= allowed[-1][None, :]).astype(np.float64))
    return masks


def masked_weights(nets, masks_obj: NumberSpinMasks, depth=2, width=64):
    """(W_abs, b_abs, W_ph, b_ph) with the MADE masks applied, from the (weight, bias) lists of make_golden.made_weights."""
    mm = made_masks(masks_obj.ends, masks_obj.Q, masks_obj.DM, depth, width)
    out = []
    for layers in nets:
        out.append([w * m for (w, _b), m in zip(layers, mm)])
        out.append([b for (_w, b) in layers])
    return out


# ---- transformer conditionals (legacy/made/real_log_psi_transformer_made.py:42-58 + the live masking, one qubit per qudit) ----
def transformer_log_psi_from_logits(logits, x, masks: NumberSpinMasks):
    """logits [B, >= n, 4] = (re0, im0, re1, im1) per position (the decoder output of TransformerMADE), x uint64 [B];
    masks built with qubit_per_qudit=1.  Returns complex128 [B]: sum over qubits of the masked, normalised conditional
    log-amplitude at the chosen outcome (ANQS:392-405 normalisation; unphysical -> -inf)."""
    n = masks.n
    x = np.asarray(x, dtype=np.uint64)
    out = np.asarray(logits)[:, :n, :].reshape(x.shape[0], n, 2, 2)
    re = np.zeros(x.shape[0])
    im = np.zeros(x.shape[0])
    rows = np.arange(x.shape[0])
    for t in range(n):
        mi = masks.memo_idx_of_prefix(x, t)
        allowed = masks.cont_mask[t][mi]                               # [B, 2]
        cond = _normalise(out[:, t, :, 0], allowed)
        bit = ((x >> np.uint64(t)) & np.uint64(1)).astype(np.int64)
        re = re + cond[rows, bit]
        im = im + np.where(allowed[rows, bit], out[rows, t, bit, 1], 0.0)
    im = np.where(np.isneginf(re), 0.0, im)
    return re + 1j * im


def transformer_cond_from_logits(logits, x, t, masks: NumberSpinMasks):
    """[B, 2] normalised conditional log|psi| of qubit t from the logits at position t."""
    x = np.asarray(x, dtype=np.uint64)
    out = np.asarray(logits)[:, t, :].reshape(x.shape[0], 2, 2)
    mi = masks.memo_idx_of_prefix(x, t)
    return _normalise(out[:, :, 0], masks.cont_mask[t][mi])


# ---- NADE mode: one MLP pair per qudit (ANQS:410-428, LAP:24-42, 63-103, 114-134) ---------------------------------------------
def nade_cond_log_abs(x_prefix, q, masks: NumberSpinMasks, W_abs_q, b_abs_q, use_res=True, subtract_mean=True):
    """[B, D_q]: W_abs_q / b_abs_q are the layer lists of log_abs_subnet[q].  The first qudit's network sees the constant
    0.5 pushed through 1 - 2x, i.e. the input 0 (MLP:205-215); the mean is over the qudit's own D_q outcomes (LAP:118-119)."""
    x_prefix = np.asarray(x_prefix, dtype=np.uint64)
    start, D = masks.starts[q], masks.dims[q]
    inp = encode(x_prefix, start) if start > 0 else np.zeros((x_prefix.shape[0], 1))
    y, _ = mlp_forward(inp, W_abs_q, b_abs_q, use_res)
    if subtract_mean:
        y = y - y.mean(axis=-1, keepdims=True)
    mi = masks.memo_idx_of_prefix(x_prefix, start)
    return _normalise(y, masks.cont_mask[q][mi])


def nade_log_psi(x, masks: NumberSpinMasks, W_abs, b_abs, W_ph, b_ph, use_res=True, subtract_mean=True):
    """complex128 [B]; W_abs[q] etc. are per-qudit layer lists."""
    x = np.asarray(x, dtype=np.uint64)
    re = np.zeros(x.shape[0])
    im = np.zeros(x.shape[0])
    rows = np.arange(x.shape[0])
    for q in range(masks.Q):
        start, D = masks.starts[q], masks.dims[q]
        cond = nade_cond_log_abs(x, q, masks, W_abs[q], b_abs[q], use_res, subtract_mean)
        inp = encode(x, start) if start > 0 else np.zeros((x.shape[0], 1))
        yp, _ = mlp_forward(inp, W_ph[q], b_ph[q], use_res)
        c = ((x >> np.uint64(start)) & np.uint64(D - 1)).astype(np.int64)
        re = re + cond[rows, c]
        im = im + np.pi * yp[rows, c]
    im = np.where(np.isneginf(re), 0.0, im)
    return re + 1j * im
