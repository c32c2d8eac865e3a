// Kernel family 4: one level of the autoregressive batch samplers.
//
//   split_level_kernel    exact multinomial split of every parent count over the D = 2^k outcomes of the next
//                         qudit as k rounds of binomial draws on a cumulative-probability tree
//                         (reference ANQS:557-591 sample_mult_new_new inside ANQS:593-662)
//   emit_children_kernel  ordered compaction of the surviving (allowed, count > 0) children (ANQS:645-660)
//   gumbel_level_kernel   conditional Gumbel perturbation of the children of every parent for stochastic-beam
//                         (top-k without replacement) sampling (ANQS:676-688, 718-731)
//
// One warp owns one parent.  The split keeps the binomial tree in registers: after round j lane t holds the
// count of tree node t (path bits most-significant first), children are handed down with two shuffles, and the
// last round leaves outcomes 2*lane and 2*lane+1 in every lane.  Random numbers are counter-based (Philox4x32-10
// keyed by the seed, counter = level / round / node / parent), so a level is reproducible independently of how
// parents are distributed over warps, launches or GPUs - which is what lets sub-trees be sharded with no
// communication.  draw_mode 0 replaces the binomial draw by its rounded mean rint(n p): the deterministic mode
// the parity tests use against the reference with torch.distributions.Binomial patched the same way.
// A node that carries a single sample takes one categorical draw by inversion instead of k binomial rounds (a multinomial
// with one trial), which is what most nodes of the deep levels of a large sparse batch are.
#include <algorithm>
#include <cstdint>

#include "common.cuh"

namespace anqs {

// ---- Philox4x32-10 ---------------------------------------------------------------------------------------
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    __device__ __forceinline__ Philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = c0; ctr[1] = c1; ctr[2] = c2; ctr[3] = c3;
    }
    __device__ __forceinline__ uint4 operator()() const {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// uniform in (0, 1) with 53 random bits
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
    uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11) ^ ((uint64_t)(b & 0x7FFu) << 42);
    v &= (1ull << 53) - 1ull;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// Binomial(n, p) variate: inversion for n*min(p,1-p) < 10, BTRS (Hormann 1993) otherwise.
__device__ double binomial_draw(double n, double p, uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2) {
    if (!(n > 0.0) || !(p > 0.0)) return 0.0;
    if (p >= 1.0) return n;
    const bool flip = p > 0.5;
    const double pp = flip ? 1.0 - p : p, q = 1.0 - pp;
    double k;
    uint32_t it = 0;
    if (n * pp < 10.0) {
        Philox g(seed, c0, c1, c2, it);
        uint4 r = g();
        double u = u01(r.x, r.y);
        const double ratio = pp / q;
        double f = exp(n * log1p(-pp));
        k = 0.0;
        while (u > f && k < n) {
            u -= f;
            k += 1.0;
            f *= ratio * (n - k + 1.0) / k;
            if (f <= 0.0) break;
        }
    } else {
        const double spq = sqrt(n * pp * q);
        const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * pp, c = n * pp + 0.5;
        const double vr = 0.92 - 4.2 / b, alpha = (2.83 + 5.1 / b) * spq;
        const double lpq = log(pp / q), m = floor((n + 1.0) * pp);
        const double h = lgamma(m + 1.0) + lgamma(n - m + 1.0);
        for (;;) {
            Philox g(seed, c0, c1, c2, it++);
            uint4 r = g();
            const double u = u01(r.x, r.y) - 0.5;
            double v = u01(r.z, r.w);
            const double us = 0.5 - fabs(u);
            k = floor((2.0 * a / us + b) * u + c);
            if (k < 0.0 || k > n) continue;
            if (us >= 0.07 && v <= vr) break;
            v = log(v * alpha / (a / (us * us) + b));
            if (v <= h - lgamma(k + 1.0) - lgamma(n - k + 1.0) + (k - m) * lpq) break;
        }
    }
    return flip ? n - k : k;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

constexpr int K4_WARPS = 8;

// cond [B][DM]: normalised conditional log|psi| (-inf = masked).  child_counts [B][D] out; n_children [B] out.
__global__ void __launch_bounds__(K4_WARPS * 32)
split_level_kernel(const double *__restrict__ cond, int DM, int k, const double *__restrict__ counts,
                   const int32_t *__restrict__ memo_idx, const unsigned long long *__restrict__ cont_mask_q,
                   int64_t memo_size, int64_t B, int level, int draw_mode, uint64_t seed, int64_t parent_offset,
                   const int64_t *__restrict__ rng_keys, double *__restrict__ child_counts, int64_t *__restrict__ n_children,
                   signed char *__restrict__ single_out, int skip_singles, int block_shift) {
    __shared__ double cum_all[K4_WARPS][66];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *cum = cum_all[warp];
    const int D = 1 << k;
    // A warp takes blocks of 32 parents, looks at their counts at once and walks the ones it has to process (all of them, or -
    // skip_singles - those that carry more than one sample: split_single_kernel took the others).  The inputs of the next
    // parent of the walk are loaded while the current one is processed (every parent starts with dependent loads: without
    // this the kernel waits on memory for half of its time).
    // blocks of 2^block_shift <= 32 parents: small levels keep one parent per warp (all warps busy), large ones 32
    const int bw = 1 << block_shift;
    const int64_t nblocks = (B + bw - 1) >> block_shift;
    double n_c0 = 0.0, n_c1 = 0.0, n_cnt = 0.0;
    int n_mi = -1;
    uint64_t n_key = 0;
    auto load_inputs = [&](int64_t bb) {
        n_c0 = lane < D ? cond[bb * DM + lane] : 0.0;
        n_c1 = lane + 32 < D ? cond[bb * DM + lane + 32] : 0.0;
        n_cnt = counts[bb];
        n_mi = memo_idx[bb];
        n_key = rng_keys ? (uint64_t)rng_keys[bb] : (uint64_t)(parent_offset + bb);
    };
    for (int64_t blk = (int64_t)blockIdx.x * K4_WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * K4_WARPS) {
    const int64_t b0 = blk << block_shift;
    const bool in_range = lane < bw && b0 + lane < B;
    const double cnt_lane = in_range ? counts[b0 + lane] : 0.0;
    unsigned todo = __ballot_sync(0xffffffffu, in_range && !(skip_singles && cnt_lane == 1.0));
    if (todo) load_inputs(b0 + (__ffs(todo) - 1));
    while (todo) {
        const int64_t b = b0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const double c0 = n_c0, c1 = n_c1, cnt_in = n_cnt;
        const int mi = n_mi;
        const uint64_t parent = n_key;
        unsigned long long mw = 0ull;
        if (mi >= 0 && mi < memo_size) mw = cont_mask_q[mi];   // needed at the end of the iteration only
        if (todo) load_inputs(b0 + (__ffs(todo) - 1));
        // probabilities = softmax(2 * logits), nan -> 0 (ANQS:560-561)
        const double l0 = lane < D ? 2.0 * c0 : -INFINITY;
        const double l1 = lane + 32 < D ? 2.0 * c1 : -INFINITY;
        double p0, p1;
        float w0 = 0.0f, w1 = 0.0f;
        if (draw_mode == 0) {
            const double mx = warp_max(fmax(l0, l1));
            double e0 = lane < D ? exp(l0 - mx) : 0.0, e1 = lane + 32 < D ? exp(l1 - mx) : 0.0;
            if (!(mx > -INFINITY)) e0 = e1 = 0.0;
            const double sum = warp_sum(e0 + e1);
            p0 = e0 / sum, p1 = e1 / sum;
            if (!(sum > 0.0)) p0 = p1 = 0.0;
            __syncwarp();
            cum[1 + lane] = p0;
            cum[33 + lane] = p1;
            __syncwarp();
        } else {
            // random draws: unnormalised weights with a single-precision exponential do (every use below is a ratio of partial
            // sums, and an unbiased relative error of 1e-7 per weight is far below anything 10^6..10^9 draws can resolve); the
            // shift is taken in double, so the weights do not inherit the rounding of the logits' magnitude, and the partial
            // sums are accumulated in double (single-precision sums fail the chi-square test at 10^7 samples: cancellation)
            float mf = (float)fmax(l0, l1);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) mf = fmaxf(mf, __shfl_xor_sync(0xffffffffu, mf, d));
            if (mf > -INFINITY) {
                w0 = lane < D ? __expf((float)(l0 - (double)mf)) : 0.0f;
                w1 = lane + 32 < D ? __expf((float)(l1 - (double)mf)) : 0.0f;
            }
            p0 = w0, p1 = w1;
        }
        double cnt = cnt_in;     // count of tree node `lane` (valid for lane < 2^j in round j)
        double c_even = 0.0, c_odd = 0.0;
        if (draw_mode == 0) {
            if (lane == 0) {  // sequential prefix sum, like a cumsum along the row (ANQS:562-565): the order the goldens were made with
                double acc = 0.0;
                cum[0] = 0.0;
                for (int d = 1; d <= D; ++d) {
                    acc += cum[d];
                    cum[d] = acc;
                }
            }
        } else {              // random draws: any summation order is as good, take the warp scan (in double: the draws below
                              // take differences of these partial sums)
            double s0 = p0, s1 = p1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
                if (lane >= d) s0 += t0, s1 += t1;
            }
            s1 += __shfl_sync(0xffffffffu, s0, 31);
            __syncwarp();
            cum[1 + lane] = s0;
            cum[33 + lane] = s1;
            if (lane == 0) cum[0] = 0.0;
        }
        __syncwarp();
        const bool single = draw_mode != 0 && cnt == 1.0;   // warp-uniform
        if (single) {
            // a multinomial with one trial is one categorical draw: invert the cumulative distribution with one uniform
            // instead of walking the binomial tree with k of them (most nodes of a deep level carry a single sample)
            Philox g(seed, (uint32_t)parent, (uint32_t)(parent >> 32), ((uint32_t)level << 16) | 0xFF00u, 0u);
            const uint4 r = g();
            const double t = u01(r.x, r.y) * cum[D];
            const bool le0 = lane + 1 <= D && cum[lane + 1] <= t, le1 = lane + 33 <= D && cum[lane + 33] <= t;
            int d = __popc(__ballot_sync(0xffffffffu, le0)) + __popc(__ballot_sync(0xffffffffu, le1));
            d = min(d, D - 1);
            // rounding at the upper end can land on an outcome of probability zero: step down to the nearest possible one
            const unsigned pos0 = __ballot_sync(0xffffffffu, p0 > 0.0), pos1 = __ballot_sync(0xffffffffu, p1 > 0.0);
            const unsigned long long pos = ((unsigned long long)pos1 << 32) | pos0;
            const unsigned long long below = pos & (d == 63 ? ~0ull : ((1ull << (d + 1)) - 1ull));
            if (below) d = 63 - __clzll(below);
            c_even = (d == 2 * lane) ? 1.0 : 0.0;
            c_odd = (d == 2 * lane + 1) ? 1.0 : 0.0;
        }
        // binomial tree, most significant outcome bit first (ANQS:568-585)
        for (int j = 0; j < (single ? 0 : k); ++j) {
            const int nodes = 1 << j, span = D >> j;
            double left = 0.0;
            if (lane < nodes) {
                const int lo = lane * span, mid = lo + (span >> 1), hi = lo + span;
                const double succ = cum[mid] - cum[lo], fail = cum[hi] - cum[mid];
                double pr = succ / (succ + fail);
                if (!(pr == pr)) pr = 0.0;  // nan_to_num (ANQS:578)
                if (draw_mode == 0) {
                    left = fmin(cnt, fmax(0.0, rint(cnt * pr)));
                } else {
                    // key of the node: its packed prefix when given (independent of how nodes are spread over launches,
                    // ranks or GPUs), else its position
                    left = binomial_draw(cnt, pr, seed, (uint32_t)parent, (uint32_t)(parent >> 32),
                                         ((uint32_t)level << 16) | ((uint32_t)j << 8) | (uint32_t)lane);
                }
            }
            if (j + 1 < k) {
                const double n_par = __shfl_sync(0xffffffffu, cnt, lane >> 1);
                const double l_par = __shfl_sync(0xffffffffu, left, lane >> 1);
                cnt = (lane & 1) ? n_par - l_par : l_par;
            } else {
                c_even = left;        // outcome 2*lane
                c_odd = cnt - left;   // outcome 2*lane + 1
            }
        }
        const int half = D >> 1;
        const bool s_even = lane < half && ((mw >> (2 * lane)) & 1ull) && c_even > 0.0;
        const bool s_odd = lane < half && ((mw >> (2 * lane + 1)) & 1ull) && c_odd > 0.0;
        const unsigned be = __ballot_sync(0xffffffffu, s_even), bo = __ballot_sync(0xffffffffu, s_odd);
        if (single_out != nullptr && single) {
            // the one child (or none, when the symmetry table forbids it) goes out as a byte: no dense row for this parent
            if (lane == 0) single_out[b] = be ? (signed char)(2 * (__ffs(be) - 1)) : bo ? (signed char)(2 * (__ffs(bo) - 1) + 1) : (signed char)-1;
        } else {
            if (lane < half) {
                child_counts[b * D + 2 * lane] = c_even;
                child_counts[b * D + 2 * lane + 1] = c_odd;
            }
            if (single_out != nullptr && lane == 0) single_out[b] = (signed char)-2;
        }
        if (lane == 0) n_children[b] = __popc(be) + __popc(bo);
    }
    }
}

// Parents that carry ONE sample - nearly all nodes of the deep levels of a large sparse batch - need one categorical draw, and a
// warp per parent spends ~450 instructions on it (shuffle reductions, a scan, ballots, all for one number).  Here a LANE owns
// a parent: a warp stages the 32 conditional rows of its parents in shared memory with coalesced loads (row stride D + 1:
// conflict-free per-lane walks), each lane then walks its own row three times - maximum, total weight, first outcome whose
// running sum passes u * total - with single-precision exponentials, the shift and the sums in double.  Same generator, same
// key (the parent's packed prefix or position) and the same inversion rule as the single-sample branch of
// split_level_kernel, so a parent's draw does not depend on which launch, rank or GPU processes it.  Parents with another
// count are left to split_level_kernel (skip_singles).
constexpr int SS_WARPS = 2;   // 33 KB of staged rows per CTA: six CTAs per SM
__global__ void __launch_bounds__(SS_WARPS * 32)
split_single_kernel(const double *__restrict__ cond, int DM, int k, const double *__restrict__ counts,
                    const int32_t *__restrict__ memo_idx, const unsigned long long *__restrict__ cont_mask_q, int64_t memo_size,
                    int64_t B, int level, uint64_t seed, int64_t parent_offset, const int64_t *__restrict__ rng_keys,
                    int64_t *__restrict__ n_children, signed char *__restrict__ single_out) {
    __shared__ double rows_all[SS_WARPS][32 * 65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *rows = rows_all[warp];
    const int D = 1 << k;
    const int64_t nblocks = (B + 31) >> 5;
    for (int64_t blk = (int64_t)blockIdx.x * SS_WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * SS_WARPS) {
        const int64_t b0 = blk << 5, b = b0 + lane;
        const bool mine = b < B && counts[b] == 1.0;
        const unsigned any = __ballot_sync(0xffffffffu, mine);
        if (!any) continue;
        // stage the rows of the block's single-sample parents: one coalesced 8 D-byte row per step, eight steps in flight
        __syncwarp();
        for (int r0 = 0; r0 < 32; r0 += 8) {
            double v0[8], v1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool want = (any >> (r0 + i)) & 1u;
                v0[i] = (want && lane < D) ? cond[(b0 + r0 + i) * DM + lane] : 0.0;
                v1[i] = (want && lane + 32 < D) ? cond[(b0 + r0 + i) * DM + lane + 32] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rows[(r0 + i) * 65 + lane] = v0[i];
                rows[(r0 + i) * 65 + 32 + lane] = v1[i];
            }
        }
        __syncwarp();
        if (mine) {
            const double *row = rows + lane * 65;
            const int mi = memo_idx[b];
            unsigned long long mw = 0ull;
            if (mi >= 0 && mi < memo_size) mw = cont_mask_q[mi];
            const uint64_t parent = rng_keys ? (uint64_t)rng_keys[b] : (uint64_t)(parent_offset + b);
            double mx = -INFINITY;
            for (int d = 0; d < D; ++d) mx = fmax(mx, row[d]);
            int pick = D - 1;
            if (mx > -INFINITY) {
                const float mf = (float)(2.0 * mx);
                double total = 0.0;
                int last_pos = -1;
                for (int d = 0; d < D; ++d) {
                    const float w = __expf((float)(2.0 * row[d] - (double)mf));   // exp(-inf) = 0 for masked outcomes
                    total += (double)w;
                    if (w > 0.0f) last_pos = d;
                }
                Philox g(seed, (uint32_t)parent, (uint32_t)(parent >> 32), ((uint32_t)level << 16) | 0xFF00u, 0u);
                const uint4 r = g();
                const double t = u01(r.x, r.y) * total;
                double cum = 0.0;
                pick = last_pos >= 0 ? last_pos : D - 1;   // rounding at the upper end: the last outcome of non-zero weight
                for (int d = 0; d < D; ++d) {
                    cum += (double)__expf((float)(2.0 * row[d] - (double)mf));
                    if (cum > t) {
                        pick = d;
                        break;
                    }
                }
            }
            const bool allowed = (mw >> pick) & 1ull;
            single_out[b] = allowed ? (signed char)pick : (signed char)-1;
            n_children[b] = allowed ? 1 : 0;
        }
    }
}

__global__ void __launch_bounds__(K4_WARPS * 32)
emit_children_kernel(const double *__restrict__ child_counts, int k, int start, const int64_t *__restrict__ prefix,
                     const int32_t *__restrict__ memo_idx, const unsigned long long *__restrict__ cont_mask_q,
                     const int32_t *__restrict__ next_memo_q, int64_t memo_size, int64_t B,
                     const int64_t *__restrict__ offsets, const signed char *__restrict__ single_in, int64_t out_cap,
                     int64_t *__restrict__ out_prefix, double *__restrict__ out_counts, int32_t *__restrict__ out_memo) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = lanemask_lt();
    const int D = 1 << k, half = D >> 1;
    // a warp looks at 32 parents at once and walks those that carry a dense row of child counts; single-sample parents (their
    // child named by one byte) are written by emit_single_kernel, a thread per parent
    const int64_t nblocks = (B + 31) >> 5;
    for (int64_t blk = (int64_t)blockIdx.x * K4_WARPS + warp; blk < nblocks; blk += (int64_t)gridDim.x * K4_WARPS) {
    const int64_t b0 = blk << 5;
    const bool in_range = b0 + lane < B;
    unsigned todo = __ballot_sync(0xffffffffu, in_range && (single_in == nullptr || single_in[b0 + lane] == -2));
    while (todo) {
        const int64_t b = b0 + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int mi = memo_idx[b];
        unsigned long long mw = 0ull;
        if (mi >= 0 && mi < memo_size) mw = cont_mask_q[mi];
        double c_even = 0.0, c_odd = 0.0;
        if (lane < half) {
            c_even = child_counts[b * D + 2 * lane];
            c_odd = child_counts[b * D + 2 * lane + 1];
        }
        const bool s_even = lane < half && ((mw >> (2 * lane)) & 1ull) && c_even > 0.0;
        const bool s_odd = lane < half && ((mw >> (2 * lane + 1)) & 1ull) && c_odd > 0.0;
        const unsigned be = __ballot_sync(0xffffffffu, s_even), bo = __ballot_sync(0xffffffffu, s_odd);
        const int64_t base = offsets[b];
        const uint64_t px = (uint64_t)prefix[b];
        const int before = __popc(be & lt) + __popc(bo & lt);
        if (s_even && base + before < out_cap) {
            const int64_t r = base + before;
            const int d = 2 * lane;
            out_prefix[r] = (int64_t)(px | ((uint64_t)d << start));
            out_counts[r] = c_even;
            out_memo[r] = next_memo_q[(size_t)mi * D + d];
        }
        if (s_odd && base + before + (s_even ? 1 : 0) < out_cap) {
            const int64_t r = base + before + (s_even ? 1 : 0);
            const int d = 2 * lane + 1;
            out_prefix[r] = (int64_t)(px | ((uint64_t)d << start));
            out_counts[r] = c_odd;
            out_memo[r] = next_memo_q[(size_t)mi * D + d];
        }
    }
    }
}

// children of the single-sample parents: one thread per parent (coalesced reads of the byte, the offset, the prefix and the
// memo index; one gathered read of the next memo index)
__global__ void __launch_bounds__(256)
emit_single_kernel(int k, int start, const int64_t *__restrict__ prefix, const int32_t *__restrict__ memo_idx,
                   const int32_t *__restrict__ next_memo_q, int64_t B, const int64_t *__restrict__ offsets,
                   const signed char *__restrict__ single_in, int64_t out_cap, int64_t *__restrict__ out_prefix,
                   double *__restrict__ out_counts, int32_t *__restrict__ out_memo) {
    const int D = 1 << k;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int sd = single_in[b];
        if (sd < 0) continue;   // -2: dense row (emit_children_kernel), -1: the one child is forbidden by the symmetries
        const int64_t r = offsets[b];
        if (r >= out_cap) continue;
        out_prefix[r] = (int64_t)((uint64_t)prefix[b] | ((uint64_t)sd << start));
        out_counts[r] = 1.0;
        out_memo[r] = next_memo_q[(size_t)memo_idx[b] * D + sd];
    }
}

__device__ __forceinline__ double log1mexp(double x) {  // ANQS:664-668
    return x > -0.693 ? log(-expm1(x)) : log1p(-exp(x));
}
__device__ __forceinline__ double log1pexp(double x) {  // ANQS:670-674
    return x < 18.0 ? log1p(exp(x)) : x + exp(-x);
}

// children of every parent: log_prob = parent log_prob + 2 cond; Gumbel conditioned on the parent's Gumbel being
// the maximum (ANQS:676-688).  Outputs are [B][D]; masked children get gumbel = -inf.
__global__ void __launch_bounds__(K4_WARPS * 32)
gumbel_level_kernel(const double *__restrict__ cond, int DM, int k, const double *__restrict__ parent_log_prob,
                    const double *__restrict__ parent_gumbel, const int32_t *__restrict__ memo_idx,
                    const unsigned long long *__restrict__ cont_mask_q, int64_t memo_size, int64_t B, int level,
                    uint64_t seed, int64_t parent_offset, const int64_t *__restrict__ rng_keys, const double *__restrict__ uniforms,
                    double *__restrict__ out_log_prob, double *__restrict__ out_gumbel) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = 1 << k;
    for (int64_t b = (int64_t)blockIdx.x * K4_WARPS + warp; b < B; b += (int64_t)gridDim.x * K4_WARPS) {
        const double lp = parent_log_prob[b], G = parent_gumbel[b];
        const int mi = memo_idx[b];
        unsigned long long mw = 0ull;
        if (mi >= 0 && mi < memo_size) mw = cont_mask_q[mi];
        double phi[2], g[2];
        bool in[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int d = lane + 32 * h;
            in[h] = d < D;
            phi[h] = -INFINITY;
            g[h] = -INFINITY;
            if (in[h]) {
                double v = lp + 2.0 * cond[b * DM + d];
                if (!(v == v)) v = -INFINITY;  // nan_to_num (ANQS:726)
                phi[h] = v;
                double u;
                if (uniforms) {
                    u = uniforms[b * D + d];
                } else {
                    // keyed by the node's own identity (its packed prefix) when given: the draws then do not depend on the order or
                    // the position of the rows of a level
                    const uint64_t parent = rng_keys ? (uint64_t)rng_keys[b] : (uint64_t)(parent_offset + b);
                    Philox rng(seed, (uint32_t)parent, (uint32_t)(parent >> 32), ((uint32_t)level << 16) | (uint32_t)d, 0x47u);
                    uint4 r = rng();
                    u = u01(r.x, r.y);
                }
                g[h] = v - log(-log(u));
            }
        }
        const double Z = warp_max(fmax(g[0], g[1]));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!in[h]) continue;
            const int d = lane + 32 * h;
            const double v = G - g[h] + log1mexp(g[h] - Z);
            double out = G - fmax(v, 0.0) - log1pexp(-fabs(v));
            if (!(out == out) || !((mw >> d) & 1ull)) out = -INFINITY;  // nan_to_num (ANQS:731); masked children never survive
            out_log_prob[b * D + d] = phi[h];
            out_gumbel[b * D + d] = out;
        }
    }
}

// ---- survivors of one Gumbel top-k level (ANQS:733-776) -------------------------------------------------------------------
// sorted_idx[r] = flat (parent * D + outcome) index of the r-th largest perturbed log-probability (the per-level global sort
// stays a library call).  Row r < keep becomes a node of the next level: prefix | outcome << qudit_start, the next memo
// index, its log-probability and its Gumbel.  Masked children carry gumbel = -inf and sort last: n_alive counts the rows
// in front of them, and the caller keeps [0, n_alive).
__global__ void __launch_bounds__(256)
gumbel_select_kernel(const int64_t *__restrict__ sorted_idx, const double *__restrict__ sorted_gumbel, int64_t keep, int k,
                     int qudit_start, const int64_t *__restrict__ prefix, const int32_t *__restrict__ memo_idx,
                     const int32_t *__restrict__ next_memo_q, const double *__restrict__ level_log_prob,
                     const unsigned long long *__restrict__ drop_mask_q, int64_t memo_size,
                     int64_t *__restrict__ out_prefix, int32_t *__restrict__ out_memo, double *__restrict__ out_log_prob,
                     double *__restrict__ out_gumbel, int *__restrict__ n_alive) {
    const int D = 1 << k;
    int alive_here = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < keep; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t flat = sorted_idx[r];
        const double g = sorted_gumbel[r];
        const int64_t parent = flat >> k;
        const int outcome = (int)(flat & (D - 1));
        bool alive = g > -INFINITY;
        if (alive && drop_mask_q != nullptr) {
            // a level drawn from UNMASKED conditionals ('DU', ANQS:708-709): unphysical children took part in the top-k with
            // finite Gumbels and are dropped only now (ANQS:804-809), wherever they sit among the kept rows
            const int mi = memo_idx[parent];
            alive = mi >= 0 && mi < memo_size && ((drop_mask_q[mi] >> outcome) & 1ull);
        }
        out_prefix[r] = prefix[parent] | ((int64_t)outcome << qudit_start);
        // dead rows (masked children, children of dead rows) may be carried to the next level: they get no memo index, which
        // masks all of their own children
        out_memo[r] = alive ? next_memo_q[(int64_t)memo_idx[parent] * D + outcome] : -1;
        out_log_prob[r] = alive ? level_log_prob[flat] : -INFINITY;
        out_gumbel[r] = alive ? g : -INFINITY;
        alive_here += alive ? 1 : 0;
    }
    alive_here = __reduce_add_sync(0xffffffffu, alive_here);
    if ((threadIdx.x & 31) == 0 && alive_here) atomicAdd(n_alive, alive_here);
}

}  // namespace anqs


using namespace anqs;

extern "C" {

int anqs_sampler_split_level(const double *d_cond, int max_qudit_dim, int qubits_in_qudit, const double *d_counts,
                             const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n,
                             int level, int draw_mode, uint64_t seed, int64_t parent_offset, const int64_t *d_rng_keys,
                             double *d_child_counts, int64_t *d_n_children, int8_t *d_single, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative parent count");
    ANQS_REQUIRE(qubits_in_qudit >= 1 && qubits_in_qudit <= 6 && (1 << qubits_in_qudit) <= max_qudit_dim && max_qudit_dim <= 64,
                 "qudit must have 1..6 qubits and fit max_qudit_dim <= 64");
    ANQS_REQUIRE(draw_mode == 0 || draw_mode == 1, "draw_mode must be 0 (rounded mean) or 1 (Philox binomial)");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_cond && d_counts && d_memo_idx && d_cont_mask_q && d_child_counts && d_n_children, "null pointer");
    // parents per warp block: as many as keeps every warp of the chip with at least one block, at most 32
    const int64_t warps = (int64_t)sm_count_of_current_device() * 8 * K4_WARPS;
    int block_shift = 0;
    while (block_shift < 5 && (n >> (block_shift + 1)) >= warps) ++block_shift;
    const int64_t nblk = (n + (1 << block_shift) - 1) >> block_shift;
    int grid = (int)std::min<int64_t>((nblk + K4_WARPS - 1) / K4_WARPS, (int64_t)sm_count_of_current_device() * 8);
    // random draws with the one-byte hand-over: single-sample parents go through the lane-per-parent kernel
    const int skip_singles = draw_mode != 0 && d_single != nullptr;
    if (skip_singles) {
        const int64_t nblocks = (n + 31) / 32;
        const int g1 = (int)std::min<int64_t>((nblocks + SS_WARPS - 1) / SS_WARPS, (int64_t)sm_count_of_current_device() * 24);
        split_single_kernel<<<g1, SS_WARPS * 32, 0, (cudaStream_t)stream>>>(
            d_cond, max_qudit_dim, qubits_in_qudit, d_counts, d_memo_idx, (const unsigned long long *)d_cont_mask_q, memo_size, n, level, seed,
            parent_offset, d_rng_keys, d_n_children, (signed char *)d_single);
        ANQS_LAUNCH_CHECK();
    }
    split_level_kernel<<<grid, K4_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d_cond, max_qudit_dim, qubits_in_qudit, d_counts, d_memo_idx, (const unsigned long long *)d_cont_mask_q, memo_size, n,
        level, draw_mode, seed, parent_offset, d_rng_keys, d_child_counts, d_n_children, (signed char *)d_single, skip_singles, block_shift);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_sampler_emit_children(const double *d_child_counts, int qubits_in_qudit, int qudit_start,
                               const int64_t *d_prefix, const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q,
                               const int32_t *d_next_memo_q, int64_t memo_size, int64_t n, const int64_t *d_offsets,
                               const int8_t *d_single, int64_t *d_out_prefix, double *d_out_counts, int32_t *d_out_memo_idx,
                               void *stream) {
    return anqs_sampler_emit_children_capped(d_child_counts, qubits_in_qudit, qudit_start, d_prefix, d_memo_idx, d_cont_mask_q, d_next_memo_q,
                                             memo_size, n, d_offsets, d_single, INT64_MAX, d_out_prefix, d_out_counts, d_out_memo_idx, stream);
}

int anqs_sampler_emit_children_capped(const double *d_child_counts, int qubits_in_qudit, int qudit_start,
                                      const int64_t *d_prefix, const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q,
                                      const int32_t *d_next_memo_q, int64_t memo_size, int64_t n, const int64_t *d_offsets,
                                      const int8_t *d_single, int64_t out_capacity, int64_t *d_out_prefix, double *d_out_counts,
                                      int32_t *d_out_memo_idx, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative parent count");
    ANQS_REQUIRE(out_capacity >= 0, "negative output capacity");
    ANQS_REQUIRE(qubits_in_qudit >= 1 && qubits_in_qudit <= 6, "qudit must have 1..6 qubits");
    ANQS_REQUIRE(qudit_start >= 0 && qudit_start + qubits_in_qudit <= 64, "qudit outside the 64-bit word");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_child_counts && d_prefix && d_memo_idx && d_cont_mask_q && d_next_memo_q && d_offsets && d_out_prefix &&
                     d_out_counts && d_out_memo_idx, "null pointer");
    const int64_t nblk = (n + 31) / 32;
    int grid = (int)std::min<int64_t>((nblk + K4_WARPS - 1) / K4_WARPS, (int64_t)sm_count_of_current_device() * 8);
    emit_children_kernel<<<grid, K4_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d_child_counts, qubits_in_qudit, qudit_start, d_prefix, d_memo_idx, (const unsigned long long *)d_cont_mask_q,
        d_next_memo_q, memo_size, n, d_offsets, (const signed char *)d_single, out_capacity, d_out_prefix, d_out_counts, d_out_memo_idx);
    ANQS_LAUNCH_CHECK();
    if (d_single != nullptr) {
        const int g1 = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count_of_current_device() * 8);
        emit_single_kernel<<<g1, 256, 0, (cudaStream_t)stream>>>(qubits_in_qudit, qudit_start, d_prefix, d_memo_idx, d_next_memo_q, n, d_offsets,
                                                                 (const signed char *)d_single, out_capacity, d_out_prefix, d_out_counts,
                                                                 d_out_memo_idx);
        ANQS_LAUNCH_CHECK();
    }
    return 0;
}

int anqs_sampler_gumbel_level(const double *d_cond, int max_qudit_dim, int qubits_in_qudit,
                              const double *d_parent_log_prob, const double *d_parent_gumbel, const int32_t *d_memo_idx,
                              const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n, int level, uint64_t seed,
                              int64_t parent_offset, const double *d_uniforms, double *d_out_log_prob,
                              double *d_out_gumbel, void *stream) {
    return anqs_sampler_gumbel_level_keyed(d_cond, max_qudit_dim, qubits_in_qudit, d_parent_log_prob, d_parent_gumbel, d_memo_idx, d_cont_mask_q,
                                           memo_size, n, level, seed, parent_offset, nullptr, d_uniforms, d_out_log_prob, d_out_gumbel, stream);
}

int anqs_sampler_gumbel_level_keyed(const double *d_cond, int max_qudit_dim, int qubits_in_qudit,
                                    const double *d_parent_log_prob, const double *d_parent_gumbel, const int32_t *d_memo_idx,
                                    const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n, int level, uint64_t seed,
                                    int64_t parent_offset, const int64_t *d_rng_keys, const double *d_uniforms, double *d_out_log_prob,
                                    double *d_out_gumbel, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative parent count");
    ANQS_REQUIRE(qubits_in_qudit >= 1 && qubits_in_qudit <= 6 && (1 << qubits_in_qudit) <= max_qudit_dim && max_qudit_dim <= 64,
                 "qudit must have 1..6 qubits and fit max_qudit_dim <= 64");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_cond && d_parent_log_prob && d_parent_gumbel && d_memo_idx && d_cont_mask_q && d_out_log_prob && d_out_gumbel,
                 "null pointer");
    int grid = (int)std::min<int64_t>((n + K4_WARPS - 1) / K4_WARPS, (int64_t)sm_count_of_current_device() * 8);
    gumbel_level_kernel<<<grid, K4_WARPS * 32, 0, (cudaStream_t)stream>>>(
        d_cond, max_qudit_dim, qubits_in_qudit, d_parent_log_prob, d_parent_gumbel, d_memo_idx,
        (const unsigned long long *)d_cont_mask_q, memo_size, n, level, seed, parent_offset, d_rng_keys, d_uniforms, d_out_log_prob,
        d_out_gumbel);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_sampler_gumbel_select(const int64_t *d_sorted_idx, const double *d_sorted_gumbel, int64_t keep, int qubits_in_qudit,
                               int qudit_start, const int64_t *d_prefix, const int32_t *d_memo_idx, const int32_t *d_next_memo_q,
                               const double *d_level_log_prob, int64_t *d_out_prefix, int32_t *d_out_memo_idx,
                               double *d_out_log_prob, double *d_out_gumbel, int32_t *d_n_alive, void *stream) {
    return anqs_sampler_gumbel_select_masked(d_sorted_idx, d_sorted_gumbel, keep, qubits_in_qudit, qudit_start, d_prefix, d_memo_idx,
                                             d_next_memo_q, d_level_log_prob, nullptr, 0, d_out_prefix, d_out_memo_idx, d_out_log_prob,
                                             d_out_gumbel, d_n_alive, stream);
}

int anqs_sampler_gumbel_select_masked(const int64_t *d_sorted_idx, const double *d_sorted_gumbel, int64_t keep, int qubits_in_qudit,
                                      int qudit_start, const int64_t *d_prefix, const int32_t *d_memo_idx,
                                      const int32_t *d_next_memo_q, const double *d_level_log_prob, const uint64_t *d_drop_mask_q,
                                      int64_t memo_size, int64_t *d_out_prefix, int32_t *d_out_memo_idx, double *d_out_log_prob,
                                      double *d_out_gumbel, int32_t *d_n_alive, void *stream) {
    ANQS_REQUIRE(keep >= 0, "negative row count");
    ANQS_REQUIRE(qubits_in_qudit >= 1 && qubits_in_qudit <= 6, "qudit must have 1..6 qubits");
    ANQS_REQUIRE(d_n_alive, "null counter");
    cudaStream_t s = (cudaStream_t)stream;
    ANQS_CUDA(cudaMemsetAsync(d_n_alive, 0, sizeof(int32_t), s));
    if (keep == 0) return 0;
    ANQS_REQUIRE(d_sorted_idx && d_sorted_gumbel && d_prefix && d_memo_idx && d_next_memo_q && d_level_log_prob && d_out_prefix &&
                     d_out_memo_idx && d_out_log_prob && d_out_gumbel, "null pointer");
    const int grid = (int)std::min<int64_t>((keep + 255) / 256, (int64_t)sm_count_of_current_device() * 8);
    gumbel_select_kernel<<<grid, 256, 0, s>>>(d_sorted_idx, d_sorted_gumbel, keep, qubits_in_qudit, qudit_start, d_prefix, d_memo_idx,
                                              d_next_memo_q, d_level_log_prob, (const unsigned long long *)d_drop_mask_q, memo_size,
                                              d_out_prefix, d_out_memo_idx, d_out_log_prob, d_out_gumbel, d_n_alive);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
