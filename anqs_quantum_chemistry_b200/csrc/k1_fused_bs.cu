// Fused sample-aware local energy, bit-sliced variant (same contract and results as fused_eloc_kernel in k1_fused.cu;
// reference PO:396-487 with coupling 'ham').
//
// The untiled kernel gives one warp one sample and spends most of its instructions on the electron-count tests (a POPC pair
// per mask, A + 0.4 U of them per sample) and on per-row bookkeeping, with ~15 % of the lanes alive at the filter probe.
// Here a warp owns a GROUP OF 32 SAMPLES:
//
//  * BIT-SLICED TESTS.  Slice i of the group = bit i of its 32 samples (one ballot).  For a sample inside the
//    (N_alpha, N_beta) sector, x' = x ^ mask keeps the electron counts iff exactly half of the positions of the alpha part
//    of the mask and half of those of its beta part are occupied in x: a boolean function of <= 8 slices that one lane
//    evaluates for 32 samples at once (~12 LOP3, no POPC).  The product layout factorises it: the alpha factor is computed
//    once per row, the beta factor once per member.
//  * SAMPLE-SYNCHRONOUS PROBES.  For every sample of the group that has a passing member in the current row step, the warp
//    probes the presence filter for that one sample: all lanes address the same 128-byte line (the line depends on the
//    sample and the row only), so the line-blocked filter keeps its one-wavefront-per-step property.  The probe address
//    and the two bit positions are XOR-linear in (sample hash, member hash): one LOP3 each per lane.
//  * Singleton rows (one member per alpha part, e.g. the same-spin double excitations) have no line to share; there each
//    lane walks the set bits of its own result word.
//  * Filter positives (true members + ~0.3 % false positives) are queued per warp as (key, mask reference, sample) and
//    resolved 32 at a time against the slot table; hits add H * psi(x') to the group's per-sample accumulators in shared
//    memory.
//  * Samples outside the sector (never produced by the symmetry-masked samplers, but legal input) take a plain path: popcount
//    tests over the flat mask table and a direct slot-table lookup per passing mask.
//
// When there are fewer groups than warps on the chip (small batches), R warps share one group and split its row steps.
#include <algorithm>

#include "common.cuh"
#include "matrix_elements.cuh"

namespace anqs {

constexpr int FB_THREADS = 1024;
constexpr int FB_WARPS = FB_THREADS / 32;
constexpr int FB_QCAP = 64;                            // queued filter positives per warp
constexpr int FB_QUEUE_BYTES = FB_WARPS * FB_QCAP * 16;  // uint4 {ka, kb, uref, sample}
constexpr uint32_t FB_UREF_ROW = 0x80000000u;          // uref flag: index into prod_row_u instead of prod_mem_u
constexpr uint32_t FB_BULK_CHUNK = 64 * 1024;
constexpr uint32_t FB_NEVER = 0x20202020u;             // four ZERO slices: the test is false for every sample

struct __align__(16) FbSlot {   // one group of 32 samples (shared memory)
    uint32_t Xa[34], Xb[34];    // slices of the alpha / beta halves; [32] = all zeros, [33] = all ones
    uint2 hs[32];               // per sample: {line hash, probe constants (offset part | bit part << 16)}
    uint32_t xa[32], xb[32];
    double acc[64];             // (re, im) of sum H psi(x') per sample: the PRIVATE accumulator of the warp with this slot's index
    uint32_t valid, slow, pad0, pad1;
};

// XOR-linear pieces of the filter address: h = hash(sample) ^ hash(member)
//   byte offset inside the (line ^ spread) group: spread bits -> line bits 7.., word bits -> 2..6
//   the key's two bits of that word, packed 5 + 5
__device__ __forceinline__ uint32_t fb_off(uint32_t h, uint32_t gmask) { return (((h >> 15) & gmask) << 7) | ((h >> 3) & 0x7Cu); }
__device__ __forceinline__ uint32_t fb_bits(uint32_t h) { return (h & 31u) | (((h >> 10) & 31u) << 5); }

__device__ __forceinline__ uint32_t fb_exactly_two(const uint32_t *X, uint32_t pos) {
    const uint32_t a = X[pos & 0xff], b = X[(pos >> 8) & 0xff], c = X[(pos >> 16) & 0xff], d = X[pos >> 24];
    return ~(a ^ b ^ c ^ d) & ~(a & b & c & d) & (a | b | c | d);
}

__device__ __forceinline__ uint32_t fb_mask_from_pos(uint32_t pos) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t p = (pos >> (8 * i)) & 0xffu;
        if (p < 32u) m |= 1u << p;
    }
    return m;
}

// both bits of the key present in the filter word
__device__ __forceinline__ uint32_t fb_test(uint32_t word, uint32_t hb) {
    return __funnelshift_r(word, 0u, hb) & __funnelshift_r(word, 0u, hb >> 5) & 1u;
}


// Slot reads of a table far larger than the L2 cache (8.4M keys: 537 MB of slots) are random DRAM sectors that are never
// reused, yet every one of them takes an L2 line from the filter words the kernel lives on: such tables are read with an
// evict-first L2 policy and no L1 allocation.  Measured on one B200, 2^20 rows (profiles/r2b_cache_hint_experiment.txt):
// 8.4M keys 18.85 -> 18.55 ms; on a 1M-key table (64 MB of slots, L2-resident and reused) the same hint costs 1 % (17.72 ->
// 17.94 ms), so it is taken from 2^22 slots on.  An evict-last policy on the filter words was also tried: +3 to +6 %, dropped.
constexpr uint32_t FB_STREAM_SLOTS_FROM = 1u << 22;
__device__ __forceinline__ long long hash_lookup_stream(const HashView &hv, uint64_t key, double &re, double &im) {
    if (key == EMPTY_KEY) return hash_lookup(hv, key, re, im);
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t h = hash_key((uint32_t)key, (uint32_t)(key >> 32)) & hv.capmask;
    for (;;) {
        unsigned long long kx, ky;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                     : "=l"(kx), "=l"(ky) : "l"(hv.slots + h), "l"(pol));
        if (kx == key) {
            double2 a = __ldg(reinterpret_cast<const double2 *>(hv.slots + h) + 1);
            re = a.x;
            im = a.y;
            return (long long)ky;
        }
        if (kx == EMPTY_KEY) return -1;
        h = (h + 1) & hv.capmask;
    }
}

// filter word of one candidate, 0 when the lane has none
__device__ __forceinline__ uint32_t fb_ldg(bool cond, const uint32_t *p) {
    uint32_t v = 0;
    if (cond) v = __ldg(p);
    return v;
}

struct FbWarp {
    FbSlot *sl;
    double *wacc;   // this warp's private accumulators (slots[warp].acc): nobody else adds to them
    uint4 *q;
    int qlen;
    const uint8_t *filter;
    uint32_t linemask, gmask;
    bool stream_slots;   // table beyond the L2 cache: slot reads with the evict-first policy (warp-uniform)
};

template <bool REAL>
__device__ __forceinline__ void fb_resolve(const Tables &t, const HashView &hv, FbWarp &w, bool active, uint4 e) {
    const uint64_t key = (uint64_t)e.x | ((uint64_t)e.y << 32);
    long long j = -1;
    double ar = 0.0, ai = 0.0;
    int2 g = make_int2(0, 0);
    if (active) {
        j = w.stream_slots ? hash_lookup_stream(hv, key, ar, ai) : hash_lookup(hv, key, ar, ai);
        if (j >= 0) {
            const uint32_t u = (e.z & FB_UREF_ROW) ? __ldg(t.prod_row_u + (e.z & ~FB_UREF_ROW)) : __ldg(t.prod_mem_u + e.z);
            g = __ldg(t.grp + u);
        }
    }
    const bool hit = active && j >= 0;
    if (__any_sync(0xffffffffu, hit)) {
        double hr, hi;
        warp_matrix_elements<REAL>(t, hit, g, key, hr, hi);
        // Deterministic accumulation: the contributions of a sample are added in queue order - across batches because the
        // batches are resolved in order into accumulators only this warp writes, inside a batch because lanes that hold the
        // same sample are summed in lane order and added by their first lane (plain adds, no atomics).
        double vr = 0.0, vi = 0.0;
        if (hit) {
            vr = REAL ? hr * ar : hr * ar - hi * ai;
            vi = REAL ? hr * ai : hr * ai + hi * ar;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, hit ? e.w : 32u + (unsigned)lane_id());
        const int nmax = __reduce_max_sync(0xffffffffu, hit ? __popc(peers) : 0);
        double *acc = w.wacc + 2 * e.w;
        if (nmax <= 1) {
            if (hit) {
                acc[0] += vr;
                acc[1] += vi;
            }
        } else {
            double sr = 0.0, si = 0.0;
            unsigned m = peers;
            for (int i = 0; i < nmax; ++i) {
                const int src = m ? __ffs(m) - 1 : lane_id();
                const double a = __shfl_sync(0xffffffffu, vr, src), b = __shfl_sync(0xffffffffu, vi, src);
                if (m) {
                    sr += a;
                    si += b;
                }
                m &= m - 1;
            }
            if (hit && lane_id() == __ffs(peers) - 1) {
                acc[0] += sr;
                acc[1] += si;
            }
        }
    }
}

// Queues the lanes flagged in `positive` (ballot b != 0); resolves a batch once 32 are waiting.
template <bool REAL>
__device__ __forceinline__ void fb_push(const Tables &t, const HashView &hv, FbWarp &w, unsigned b, bool positive, uint32_t s,
                                        uint32_t apos, uint32_t bpos, uint32_t uref) {
    const int lane = lane_id();
    if (positive) {
        const int p = w.qlen + __popc(b & lanemask_lt());
        const uint32_t ka = w.sl->xa[s] ^ fb_mask_from_pos(apos), kb = w.sl->xb[s] ^ fb_mask_from_pos(bpos);
        w.q[p] = make_uint4(ka, kb, uref, s);
    }
    w.qlen += __popc(b);
    __syncwarp();
    if (w.qlen >= 32) {
        const uint4 e0 = w.q[lane];
        const int rem = w.qlen - 32;
        uint4 e1 = make_uint4(0, 0, 0, 0);
        if (lane < rem) e1 = w.q[32 + lane];
        __syncwarp();
        if (lane < rem) w.q[lane] = e1;
        __syncwarp();
        w.qlen = rem;
        fb_resolve<REAL>(t, hv, w, true, e0);
    }
}

template <bool REAL>
__device__ __forceinline__ void fb_process_tile(const Tables &t, const HashView &hv, FbWarp &w, const ProdTile &tile,
                                                const unsigned char *smem_tile, int k, int R) {
    const int lane = lane_id();
    const FbSlot &sl = *w.sl;
    const uint32_t valid = sl.valid;
    // per-sample probe constants through a 32-bit shared-space address (a generic pointer costs a window conversion per use)
    const uint32_t hs_addr = (uint32_t)__cvta_generic_to_shared(sl.hs);
    const uint4 *rows = reinterpret_cast<const uint4 *>(smem_tile);
    const uint2 *mems = reinterpret_cast<const uint2 *>(smem_tile + (size_t)(tile.n_multi + tile.n_single) * sizeof(RowRec));
    // ---- multi-member rows: alpha factor per row, beta factor per member, one probe step per (row step, passing sample) ----
    // warp k of the group takes the steps j of row r with (r + j) % R == k; jrot = (k - r) mod R, kept without divisions
    uint32_t jrot = (uint32_t)k;
    for (uint32_t r = 0; r < tile.n_multi; ++r, jrot = jrot ? jrot - 1u : (uint32_t)R - 1u) {
        const uint4 rec = rows[r];  // {alpha positions, line hash of the alpha part, first member, member count}
        const uint32_t nsteps = (rec.w + 63u) >> 6;
        uint32_t j = jrot;
        if (j >= nsteps) continue;
        const uint32_t fa = fb_exactly_two(sl.Xa, rec.x) & valid;
        if (!fa) continue;
        for (; j < nsteps; j += R) {
            const uint32_t j1 = j * 64u + lane, j2 = j1 + 32u;
            uint2 m1 = make_uint2(FB_NEVER, 0u), m2 = make_uint2(FB_NEVER, 0u);
            if (j1 < rec.w) m1 = mems[rec.z + j1];
            if (j2 < rec.w) m2 = mems[rec.z + j2];
            const uint32_t w1 = fa & fb_exactly_two(sl.Xb, m1.x), w2 = fa & fb_exactly_two(sl.Xb, m2.x);
            uint32_t any = __reduce_or_sync(0xffffffffu, w1 | w2);
            if (!any) continue;
            const uint32_t mo1 = fb_off(m1.y, w.gmask), mo2 = fb_off(m2.y, w.gmask), mh1 = fb_bits(m1.y), mh2 = fb_bits(m2.y);
            while (any) {
                // two samples per iteration: four independent filter loads in flight per lane
                const uint32_t sa = __ffs(any) - 1;
                any &= any - 1;
                const bool two = any != 0u;
                const uint32_t sb = two ? __ffs(any) - 1 : sa;
                any &= any - 1;  // 0 stays 0
                uint2 ha, hb;  // uniform addresses
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ha.x), "=r"(ha.y) : "r"(hs_addr + sa * 8u));
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hb.x), "=r"(hb.y) : "r"(hs_addr + sb * 8u));
                const uint32_t ua = (((ha.x ^ rec.y) & w.linemask) << 7) ^ (ha.y & 0xFFFFu);
                const uint32_t ub = (((hb.x ^ rec.y) & w.linemask) << 7) ^ (hb.y & 0xFFFFu);
                const uint32_t bita = 1u << sa, bitb = two ? 1u << sb : 0u;
                uint32_t wa1, wa2, wb1, wb2;
                // filter is FILTER_ALIGN-aligned and mo < FILTER_ALIGN: (filter + u) ^ mo == filter + (u ^ mo), one LOP3 per probe
                const uintptr_t pa = (uintptr_t)(w.filter + ua), pb = (uintptr_t)(w.filter + ub);
                wa1 = fb_ldg(w1 & bita, reinterpret_cast<const uint32_t *>(pa ^ mo1));
                wa2 = fb_ldg(w2 & bita, reinterpret_cast<const uint32_t *>(pa ^ mo2));
                wb1 = fb_ldg(w1 & bitb, reinterpret_cast<const uint32_t *>(pb ^ mo1));
                wb2 = fb_ldg(w2 & bitb, reinterpret_cast<const uint32_t *>(pb ^ mo2));
                const uint32_t ka = ha.y >> 16, kb = hb.y >> 16;
                const bool fa1 = fb_test(wa1, ka ^ mh1), fa2 = fb_test(wa2, ka ^ mh2);
                const bool fb1 = fb_test(wb1, kb ^ mh1), fb2 = fb_test(wb2, kb ^ mh2);
                if (__any_sync(0xffffffffu, fa1 | fa2 | fb1 | fb2)) {
                    unsigned b = __ballot_sync(0xffffffffu, fa1);
                    if (b) fb_push<REAL>(t, hv, w, b, fa1, sa, rec.x, m1.x, tile.member_base + rec.z + j1);
                    b = __ballot_sync(0xffffffffu, fa2);
                    if (b) fb_push<REAL>(t, hv, w, b, fa2, sa, rec.x, m2.x, tile.member_base + rec.z + j2);
                    b = __ballot_sync(0xffffffffu, fb1);
                    if (b) fb_push<REAL>(t, hv, w, b, fb1, sb, rec.x, m1.x, tile.member_base + rec.z + j1);
                    b = __ballot_sync(0xffffffffu, fb2);
                    if (b) fb_push<REAL>(t, hv, w, b, fb2, sb, rec.x, m2.x, tile.member_base + rec.z + j2);
                }
            }
        }
    }
    // ---- singleton rows: one row per lane, every lane walks the set bits (samples) of its own result word --------------
    const uint4 *srows = rows + tile.n_multi;
    for (uint32_t r0 = (uint32_t)k * 32u; r0 < tile.n_single; r0 += 32u * (uint32_t)R) {
        const uint32_t r = r0 + lane;
        uint4 c = make_uint4(FB_NEVER, 0u, FB_NEVER, 0u);  // {alpha positions, line hash, beta positions, member hash}
        if (r < tile.n_single) c = srows[r];
        uint32_t ww = valid & fb_exactly_two(sl.Xa, c.x) & fb_exactly_two(sl.Xb, c.z);
        const uint32_t mo = fb_off(c.w, w.gmask), mh = fb_bits(c.w);
        while (__any_sync(0xffffffffu, ww != 0u)) {
            // two of the lane's own samples per iteration
            const bool acta = ww != 0u;
            const uint32_t sa = acta ? __ffs(ww) - 1 : 0u;
            ww &= ww - 1;  // 0 stays 0
            const bool actb = ww != 0u;
            const uint32_t sb = actb ? __ffs(ww) - 1 : 0u;
            ww &= ww - 1;
            const uint2 ha = sl.hs[sa], hb = sl.hs[sb];
            const uint32_t offa = ((((ha.x ^ c.y) & w.linemask) << 7) ^ (ha.y & 0xFFFFu)) ^ mo;
            const uint32_t offb = ((((hb.x ^ c.y) & w.linemask) << 7) ^ (hb.y & 0xFFFFu)) ^ mo;
            const uint32_t worda = fb_ldg(acta, reinterpret_cast<const uint32_t *>(w.filter + offa));
            const uint32_t wordb = fb_ldg(actb, reinterpret_cast<const uint32_t *>(w.filter + offb));
            const bool fa = fb_test(worda, (ha.y >> 16) ^ mh), fb = fb_test(wordb, (hb.y >> 16) ^ mh);
            if (__any_sync(0xffffffffu, fa | fb)) {
                unsigned b = __ballot_sync(0xffffffffu, fa);
                if (b) fb_push<REAL>(t, hv, w, b, fa, sa, c.x, c.z, FB_UREF_ROW | (tile.row_base + tile.n_multi + r));
                b = __ballot_sync(0xffffffffu, fb);
                if (b) fb_push<REAL>(t, hv, w, b, fb, sb, c.x, c.z, FB_UREF_ROW | (tile.row_base + tile.n_multi + r));
            }
        }
    }
}

// Plain path for one sample outside the sector: popcount tests over the flat mask table, direct slot lookups.
template <bool REAL>
__device__ __forceinline__ void fb_slow_sample(const Tables &t, const HashView &hv, uint32_t xa, uint32_t xb, int alpha, int beta,
                                               double &er, double &ei) {
    const int lane = lane_id();
    double sr = 0.0, si = 0.0;
    for (int64_t u0 = 0; u0 < t.U; u0 += 32) {
        const int64_t u = u0 + lane;
        uint2 m = make_uint2(0u, 0u);
        if (u < t.U) m = __ldg(t.mab + u);
        const bool pass = u < t.U && __popc(xa ^ m.x) == alpha && __popc(xb ^ m.y) == beta;
        const uint64_t key = (uint64_t)(xa ^ m.x) | ((uint64_t)(xb ^ m.y) << 32);
        long long j = -1;
        double ar = 0.0, ai = 0.0;
        if (pass) j = hash_lookup(hv, key, ar, ai);
        const bool hit = pass && j >= 0;
        if (__any_sync(0xffffffffu, hit)) {
            int2 g = make_int2(0, 0);
            if (hit) g = __ldg(t.grp + u);
            double hr, hi;
            warp_matrix_elements<REAL>(t, hit, g, key, hr, hi);
            if (hit) {
                if (REAL) {
                    sr += hr * ar;
                    si += hr * ai;
                } else {
                    sr += hr * ar - hi * ai;
                    si += hr * ai + hi * ar;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, d);
        si += __shfl_xor_sync(0xffffffffu, si, d);
    }
    er = sr;
    ei = si;
}

// SPLIT = false: every warp owns a group (S = FB_WARPS, R = 1 at compile time: the shape of large batches, kept free of the
// bookkeeping of the general case); SPLIT = true: S groups per CTA iteration, R warps per group, both run-time.
template <bool REAL, bool SPLIT>
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_eloc_bs_kernel(Tables t, HashView hv, const int64_t *__restrict__ samples, const double2 *__restrict__ amps,
                     int64_t row_start, int64_t row_len, int alpha, int beta, double2 *__restrict__ eloc, int S_arg, int R_arg) {
    const int S = SPLIT ? S_arg : FB_WARPS, R = SPLIT ? R_arg : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ ProdTile s_tile;
    FbSlot *slots = reinterpret_cast<FbSlot *>(smem_raw);
    uint4 *queues = reinterpret_cast<uint4 *>(smem_raw + FB_WARPS * sizeof(FbSlot));
    unsigned char *tile_buf = smem_raw + FB_WARPS * sizeof(FbSlot) + FB_QUEUE_BYTES;

    const int warp = threadIdx.x >> 5, lane = lane_id();
    // S groups per CTA iteration, R warps per group (S * R <= FB_WARPS; the warps left over only take part in the barriers)
    const bool live = SPLIT ? warp < S * R : true;
    const int slot = SPLIT ? (live ? warp / R : 0) : warp, k = SPLIT ? warp - (warp / R) * R : 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;
    FbWarp w;
    w.sl = slots + slot;
    w.wacc = slots[warp].acc;
    w.q = queues + warp * FB_QCAP;
    w.filter = hv.filter;
    w.linemask = hv.linemask;
    w.gmask = __ldg(&hv.header->gmask);
    w.stream_slots = hv.capmask >= FB_STREAM_SLOTS_FROM - 1u;
    FbSlot &sl = *w.sl;

    const bool resident = t.n_tiles == 1;
    bool loaded = false;
    const int64_t ngroups = (row_len + 31) >> 5;
    const int64_t per_iter = (int64_t)gridDim.x * S;
    for (int64_t g0 = (int64_t)blockIdx.x * S; g0 < ngroups; g0 += per_iter) {
        const int64_t g = g0 + slot;
        const bool have = live && g < ngroups;
        __syncthreads();  // the previous groups are finished with the slots
        if (k == 0 && live) {
            const int64_t r = g * 32 + lane;
            const bool ok = have && r < row_len;
            const uint64_t x = ok ? (uint64_t)samples[row_start + r] : 0ull;
            const uint32_t xa = compress_even_bits(x), xb = compress_even_bits(x >> 1);
            const bool insec = __popc(xa) == alpha && __popc(xb) == beta;
            uint32_t ma = 0, mb = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t ba = __ballot_sync(0xffffffffu, (xa >> i) & 1u), bb = __ballot_sync(0xffffffffu, (xb >> i) & 1u);
                if (lane == i) {
                    ma = ba;
                    mb = bb;
                }
            }
            sl.Xa[lane] = ma;
            sl.Xb[lane] = mb;
            if (lane < 2) {
                sl.Xa[32 + lane] = lane ? 0xffffffffu : 0u;
                sl.Xb[32 + lane] = lane ? 0xffffffffu : 0u;
            }
            const uint32_t hl = lin_dev(LIN_LINE, xa);
            const uint32_t hp = (lin_dev(LIN_POSA, xa) & POSA_MASK) ^ (lin_dev(LIN_POSB, xb) & POSB_MASK);
            sl.hs[lane] = make_uint2(hl, fb_off(hp, w.gmask) | (fb_bits(hp) << 16));
            sl.xa[lane] = xa;
            sl.xb[lane] = xb;
            const uint32_t valid = __ballot_sync(0xffffffffu, ok && insec), slow = __ballot_sync(0xffffffffu, ok && !insec);
            if (lane == 0) {
                sl.valid = valid;
                sl.slow = slow;
            }
        }
        w.wacc[2 * lane] = 0.0;   // every warp clears its own accumulators (the previous groups' sums were taken before the barrier above)
        w.wacc[2 * lane + 1] = 0.0;
        __syncthreads();
        w.qlen = 0;
        for (int ti = 0; ti < t.n_tiles; ++ti) {
            if (!resident || !loaded) {
                __syncthreads();  // everyone is done with the previous contents of tile_buf / s_tile
                if (threadIdx.x == 0) {
                    const ProdTile pt = t.prod_tiles[ti];
                    s_tile = pt;
                    mbar_arrive_expect_tx(&bar, pt.blob_bytes);
                    for (uint32_t off = 0; off < pt.blob_bytes; off += FB_BULK_CHUNK)
                        bulk_copy_g2s(tile_buf + off, t.prod_blob_bs + pt.blob_off + off, min(FB_BULK_CHUNK, pt.blob_bytes - off), &bar);
                }
                __syncthreads();
                mbar_wait(&bar, parity);
                parity ^= 1u;
                loaded = true;
            }
            const ProdTile tile = s_tile;
            if (have && sl.valid) fb_process_tile<REAL>(t, hv, w, tile, tile_buf, k, R);
        }
        __syncwarp();
        if (w.qlen > 0) {
            const bool act = lane < w.qlen;
            const uint4 e = act ? w.q[lane] : make_uint4(0, 0, 0, 0);
            fb_resolve<REAL>(t, hv, w, act, e);
            w.qlen = 0;
        }
        __syncthreads();  // every contribution to the accumulators of this CTA's groups has landed
        if (k == 0 && have) {
            // the group's sums: the private accumulators of its R warps, added in warp order
            double sr = 0.0, si = 0.0;
            for (int kk = 0; kk < R; ++kk) {
                sr += slots[slot * R + kk].acc[2 * lane];
                si += slots[slot * R + kk].acc[2 * lane + 1];
            }
            uint32_t slow = sl.slow;
            while (slow) {
                const uint32_t s = __ffs(slow) - 1;
                slow &= slow - 1;
                double er, ei;
                fb_slow_sample<REAL>(t, hv, sl.xa[s], sl.xb[s], alpha, beta, er, ei);
                if (lane == (int)s) {
                    sr = er;
                    si = ei;
                }
            }
            if (((sl.valid | sl.slow) >> lane) & 1u) {
                const int64_t r = g * 32 + lane;
                const double2 a = amps[row_start + r];
                const double den = a.x * a.x + a.y * a.y;
                eloc[r] = make_double2((sr * a.x + si * a.y) / den, (si * a.x - sr * a.y) / den);
            }
        }
    }
}

// returns 1 when it launched, 0 when the bit-sliced kernel does not apply (the caller falls back), < 0 on error
// variant: 0 = chosen by batch size, 1 = never (the caller's warp-per-sample kernel), 2 = always when the table allows it
int fused_bs_try_launch(const Tables *t, HashView hv, const int64_t *d_samples, const double *d_amps, int64_t row_start,
                        int64_t row_len, int alpha_num, int beta_num, double *d_eloc, int variant, cudaStream_t s) {
    if (variant == 1 || !t->prod_bs_ok) return 0;
    const size_t smem = (size_t)FB_WARPS * sizeof(FbSlot) + FB_QUEUE_BYTES + (size_t)t->tile_bytes_max;
    if (smem + 256 > 227 * 1024) return 0;
    const int sms = sm_count_of_current_device();
    const int64_t ngroups = (row_len + 31) / 32;
    // small batches leave most SMs without a group: the warp-per-sample kernel spreads them better
    if (variant == 0 && ngroups < (int64_t)sms * 8) return 0;
    // S groups per CTA iteration, R = FB_WARPS / S warps per group: as few groups per CTA as spreads them over all the SMs
    const int S = (int)std::min<int64_t>(FB_WARPS, (ngroups + sms - 1) / sms);
    const int R = FB_WARPS / S;
    const int grid = (int)std::min<int64_t>((ngroups + S - 1) / S, sms);
    cudaError_t e;
    // Presence filters beyond ~24 MB fall out of L2 under their own single-sector random traffic (profiles/README.md):
    // ask for the filter to be kept as persisting L2 lines for this launch.  Best effort - every call may be refused.
    const size_t filter_bytes = ((size_t)hv.linemask + 1) * 128;
    bool window = false;
    cudaStreamAttrValue prev_attr = {};
    if (filter_bytes > ((size_t)16 << 20)) {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (max_persist > 0 && max_window > 0) {
            const size_t want = std::min<size_t>(filter_bytes, (size_t)max_persist);
            static size_t limit_set[64] = {0};  // per device: the set-aside only ever grows, and is not touched again once large enough
            if (dev >= 0 && dev < 64 && limit_set[dev] < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
                limit_set[dev] = want;
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = const_cast<uint8_t *>(hv.filter);
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(filter_bytes, (size_t)max_window);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)filter_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            // the caller's own window (if any) is put back after the launch
            if (cudaStreamGetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &prev_attr) != cudaSuccess) prev_attr = cudaStreamAttrValue{};
            window = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
        }
        cudaGetLastError();  // refusals are not errors of this call
    }
    auto kern = t->weights_real ? (S == FB_WARPS ? fused_eloc_bs_kernel<true, false> : fused_eloc_bs_kernel<true, true>)
                                : (S == FB_WARPS ? fused_eloc_bs_kernel<false, false> : fused_eloc_bs_kernel<false, true>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -1;
    kern<<<grid, FB_THREADS, smem, s>>>(*t, hv, d_samples, (const double2 *)d_amps, row_start, row_len, alpha_num, beta_num,
                                        (double2 *)d_eloc, S, R);
    if (window) {  // the window applies to the launches issued while it is set: restore what the stream had before
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &prev_attr);
        cudaGetLastError();
    }
    return 1;
}

}  // namespace anqs

