// Fused sample-aware local energy (reference PO:396-487 compute_var_local_energy_proxy with coupling 'ham',
// i.e. PO:527-567 candidates + filter, HS:263-284 join, PO:256-324 matrix elements, PO:453-478 accumulate):
//
//   E_loc[i] = ( sum over masks u with x' = x_i ^ xy[u] physical and x' in the sampled set of  H_{x_i,x'} psi(x') ) / psi(x_i)
//
// Nothing is materialised.  Design (all of it is about not doing U tests and U random loads per sample):
//
//  * PRODUCT LAYOUT.  Every XY mask is (alpha part pa, beta part mb) after de-interleaving, and the electron-count
//    filter factorises: popc(xa ^ pa) == N_alpha and popc(xb ^ mb) == N_beta.  Masks are grouped into rows by their
//    alpha part (Tables::prod_*, built in abi_core.cu): a warp tests the ~A distinct alpha parts once per sample and
//    sweeps only the members of the rows that pass, ~0.4 U beta tests instead of 2 U tests.
//  * LINEAR HASHES.  The presence filter in front of the lookup table (k2_hash.cu) is addressed by GF(2)-linear
//    hashes, so the filter address of x' = x ^ mask is hash(x) ^ hash(mask): one XOR with a constant stored in the
//    member record, no per-candidate mixing.
//  * LINE-BLOCKED FILTER.  The 128-byte filter line depends only on the alpha half of the key, so every candidate of
//    one (sample, row) pair tests a bit of the same line: one L1 wavefront per warp step.
//  * Candidates that pass the filter bit (true members + ~2 % false positives) are queued per warp and resolved 32
//    at a time against the slot table; hits get their matrix element and are accumulated in fp64.
//
// One warp per sample, 32 warps per CTA, one persistent CTA per SM.  The product tiles are staged into shared memory
// by 1-D bulk TMA copies (cp.async.bulk + mbarrier): resident when everything fits one tile (<= 200 KB), otherwise
// re-streamed tile by tile for every group of 32 samples.
#include <algorithm>

#include "common.cuh"
#include "matrix_elements.cuh"

namespace anqs {

constexpr int FZ_THREADS = 1024;
constexpr int FZ_WARPS = FZ_THREADS / 32;
constexpr int FZ_QCAP = 64;                                 // queued filter positives per warp
constexpr int FZ_QUEUE_BYTES = FZ_WARPS * FZ_QCAP * 3 * 4;  // (ka, kb, uref) per entry
constexpr uint32_t UREF_ROW = 0x80000000u;                  // uref flag: index into prod_row_u instead of prod_mem_u
constexpr uint32_t BULK_CHUNK = 64 * 1024;                  // bytes per bulk copy

struct FzWarp {
    // sample
    uint32_t xa, xb;
    int alpha, beta;
    uint32_t hl, hp;          // linear hashes of the sample
    // filter
    const uint8_t *filter;
    uint32_t linemask, gshift;  // gshift = gmask << 7
    // positives queue (shared memory, per warp)
    uint32_t *q_ka, *q_kb, *q_u;
    int qlen;
    // accumulators (per lane partial sums)
    double er, ei;
};

// Resolves up to 32 queued candidates against the slot table; hits get H_{x,x'} * psi(x') accumulated.
template <bool REAL>
__device__ __forceinline__ void fz_resolve(const Tables &t, const HashView &hv, FzWarp &w, bool active, uint32_t ka,
                                           uint32_t kb, uint32_t uref) {
    uint64_t key = (uint64_t)ka | ((uint64_t)kb << 32);
    long long j = -1;
    double ar = 0.0, ai = 0.0;
    int2 g = make_int2(0, 0);
    if (active) {
        j = hash_lookup(hv, key, ar, ai);
        if (j >= 0) {
            const uint32_t u = (uref & UREF_ROW) ? __ldg(t.prod_row_u + (uref & ~UREF_ROW)) : __ldg(t.prod_mem_u + uref);
            g = __ldg(t.grp + u);
        }
    }
    const bool hit = active && j >= 0;
    if (__any_sync(0xffffffffu, hit)) {
        double hr, hi;
        warp_matrix_elements<REAL>(t, hit, g, key, hr, hi);
        if (hit) {
            if (REAL) {
                w.er += hr * ar;
                w.ei += hr * ai;
            } else {
                w.er += hr * ar - hi * ai;
                w.ei += hr * ai + hi * ar;
            }
        }
    }
}

// Appends the candidates flagged in `positive` to the warp's queue; resolves a batch when 32 are waiting.
// `b` is the ballot of `positive` (non-zero).
template <bool REAL>
__device__ __forceinline__ void fz_push(const Tables &t, const HashView &hv, FzWarp &w, unsigned b, bool positive,
                                        uint32_t ka, uint32_t kb, uint32_t uref) {
    const int lane = lane_id();
    if (positive) {
        const int p = w.qlen + __popc(b & lanemask_lt());
        w.q_ka[p] = ka;
        w.q_kb[p] = kb;
        w.q_u[p] = uref;
    }
    w.qlen += __popc(b);
    __syncwarp();
    if (w.qlen >= 32) {
        const uint32_t a0 = w.q_ka[lane], b0 = w.q_kb[lane], u0 = w.q_u[lane];
        const int rem = w.qlen - 32;
        uint32_t a1 = 0, b1 = 0, u1 = 0;
        if (lane < rem) {
            a1 = w.q_ka[32 + lane];
            b1 = w.q_kb[32 + lane];
            u1 = w.q_u[32 + lane];
        }
        __syncwarp();
        if (lane < rem) {
            w.q_ka[lane] = a1;
            w.q_kb[lane] = b1;
            w.q_u[lane] = u1;
        }
        __syncwarp();
        w.qlen = rem;
        fz_resolve<REAL>(t, hv, w, true, a0, b0, u0);
    }
}

// Two probe results per lane (one double step): a single vote decides whether anything has to be queued at all.
template <bool REAL>
__device__ __forceinline__ void fz_push2(const Tables &t, const HashView &hv, FzWarp &w, bool f1, bool f2, uint32_t ka1,
                                         uint32_t kb1, uint32_t u1, uint32_t ka2, uint32_t kb2, uint32_t u2) {
    if (!__any_sync(0xffffffffu, f1 | f2)) return;
    const unsigned b1 = __ballot_sync(0xffffffffu, f1);
    if (b1) fz_push<REAL>(t, hv, w, b1, f1, ka1, kb1, u1);
    const unsigned b2 = __ballot_sync(0xffffffffu, f2);
    if (b2) fz_push<REAL>(t, hv, w, b2, f2, ka2, kb2, u2);
}

// filter bit of the candidate with member hash `mhash` in the line group `rowline` (byte offset of the row's line)
__device__ __forceinline__ bool fz_filter_bit(const FzWarp &w, bool pass, uint32_t rowline, uint32_t mhash) {
    const uint32_t h = w.hp ^ mhash;
    const uint32_t off = (rowline ^ ((h >> 8) & w.gshift)) + ((h >> 3) & 0x7Cu);  // line * 128 + word * 4
    const uint32_t pat = (1u << (h & 31u)) | (1u << ((h >> 10) & 31u));            // the key's two bits of that word
    uint32_t word = 0;
    if (pass) word = __ldg(reinterpret_cast<const uint32_t *>(w.filter + off));  // 32-bit offset: filters are <= 1 GiB
    return (word & pat) == pat;
}

template <bool REAL>
__device__ __forceinline__ void fz_process_tile(const Tables &t, const HashView &hv, FzWarp &w, const ProdTile &tile,
                                                const unsigned char *smem_tile) {
    const int lane = lane_id();
    const uint4 *rows = reinterpret_cast<const uint4 *>(smem_tile);
    const uint2 *mems = reinterpret_cast<const uint2 *>(smem_tile + (size_t)(tile.n_multi + tile.n_single) * sizeof(RowRec));
    // ---- multi-member rows: alpha test on 32 rows per step, then the members of every passing row ----------
    for (uint32_t r0 = 0; r0 < tile.n_multi; r0 += 32) {
        const uint32_t r = r0 + lane;
        bool pass_a = false;
        if (r < tile.n_multi) pass_a = __popc(w.xa ^ rows[r].x) == w.alpha;
        unsigned todo = __ballot_sync(0xffffffffu, pass_a);
        while (todo) {
            const uint32_t rr = r0 + (__ffs(todo) - 1);
            todo &= todo - 1;
            const uint4 rec = rows[rr];  // uniform address: one broadcast load
            const uint32_t ka = w.xa ^ rec.x;
            const uint32_t rowline = ((w.hl ^ rec.y) & w.linemask) << 7;
            const uint32_t start = rec.z, len = rec.w;
            for (uint32_t j0 = 0; j0 < len; j0 += 64) {
                const uint32_t j1 = j0 + lane, j2 = j0 + 32 + lane;
                uint2 m1 = make_uint2(0, 0), m2 = make_uint2(0, 0);
                if (j1 < len) m1 = mems[start + j1];
                if (j2 < len) m2 = mems[start + j2];
                const bool p1 = j1 < len && __popc(w.xb ^ m1.x) == w.beta;
                const bool p2 = j2 < len && __popc(w.xb ^ m2.x) == w.beta;
                const bool f1 = fz_filter_bit(w, p1, rowline, m1.y);
                const bool f2 = fz_filter_bit(w, p2, rowline, m2.y);
                fz_push2<REAL>(t, hv, w, f1, f2, ka, w.xb ^ m1.x, tile.member_base + start + j1, ka, w.xb ^ m2.x,
                               tile.member_base + start + j2);
            }
        }
    }
    // ---- singleton rows: both tests and the filter probe in the lane that owns the row -----------------------
    const uint4 *srows = rows + tile.n_multi;
    for (uint32_t r0 = 0; r0 < tile.n_single; r0 += 64) {
        const uint32_t r1 = r0 + lane, r2 = r0 + 32 + lane;
        uint4 c1 = make_uint4(0, 0, 0, 0), c2 = make_uint4(0, 0, 0, 0);
        if (r1 < tile.n_single) c1 = srows[r1];
        if (r2 < tile.n_single) c2 = srows[r2];
        const bool p1 = r1 < tile.n_single && __popc(w.xa ^ c1.x) == w.alpha && __popc(w.xb ^ c1.z) == w.beta;
        const bool p2 = r2 < tile.n_single && __popc(w.xa ^ c2.x) == w.alpha && __popc(w.xb ^ c2.z) == w.beta;
        const bool f1 = fz_filter_bit(w, p1, ((w.hl ^ c1.y) & w.linemask) << 7, c1.w);
        const bool f2 = fz_filter_bit(w, p2, ((w.hl ^ c2.y) & w.linemask) << 7, c2.w);
        fz_push2<REAL>(t, hv, w, f1, f2, w.xa ^ c1.x, w.xb ^ c1.z, UREF_ROW | (tile.row_base + tile.n_multi + r1), w.xa ^ c2.x,
                       w.xb ^ c2.z, UREF_ROW | (tile.row_base + tile.n_multi + r2));
    }
}

template <bool REAL>
__global__ void __launch_bounds__(FZ_THREADS, 1)
fused_eloc_kernel(Tables t, HashView hv, const int64_t *__restrict__ samples, const double2 *__restrict__ amps,
                  int64_t row_start, int64_t row_len, int alpha, int beta, double2 *__restrict__ eloc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ ProdTile s_tile;
    uint32_t *queues = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned char *tile_buf = smem_raw + FZ_QUEUE_BYTES;

    const int warp = threadIdx.x >> 5, lane = lane_id();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    FzWarp w;
    w.filter = hv.filter;
    w.linemask = hv.linemask;
    w.gshift = __ldg(&hv.header->gmask) << 7;
    w.q_ka = queues + warp * (FZ_QCAP * 3);
    w.q_kb = w.q_ka + FZ_QCAP;
    w.q_u = w.q_kb + FZ_QCAP;
    w.beta = beta;

    const bool resident = t.n_tiles == 1;
    bool loaded = false;
    const int64_t ngroups = (row_len + FZ_WARPS - 1) / FZ_WARPS;
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
        const int64_t r = group * FZ_WARPS + warp;
        const bool have = r < row_len;
        const uint64_t x = have ? (uint64_t)samples[row_start + r] : 0ull;
        w.xa = compress_even_bits(x);
        w.xb = compress_even_bits(x >> 1);
        w.alpha = have ? alpha : -1;  // rows past the end get an impossible electron count: nothing passes
        w.hl = lin_warp(LIN_LINE, w.xa);
        w.hp = (lin_warp(LIN_POSA, w.xa) & POSA_MASK) ^ (lin_warp(LIN_POSB, w.xb) & POSB_MASK);
        w.qlen = 0;
        w.er = w.ei = 0.0;
        for (int ti = 0; ti < t.n_tiles; ++ti) {
            if (!resident || !loaded) {
                __syncthreads();  // everyone is done with the previous contents of tile_buf / s_tile
                if (threadIdx.x == 0) {
                    const ProdTile pt = t.prod_tiles[ti];
                    s_tile = pt;
                    mbar_arrive_expect_tx(&bar, pt.blob_bytes);
                    for (uint32_t off = 0; off < pt.blob_bytes; off += BULK_CHUNK)
                        bulk_copy_g2s(tile_buf + off, t.prod_blob + pt.blob_off + off, min(BULK_CHUNK, pt.blob_bytes - off), &bar);
                }
                __syncthreads();  // s_tile visible
                mbar_wait(&bar, parity);
                parity ^= 1u;
                loaded = true;
            }
            const ProdTile tile = s_tile;
            if (have) fz_process_tile<REAL>(t, hv, w, tile, tile_buf);
        }
        if (have) {
            __syncwarp();
            if (w.qlen > 0) {
                const bool act = lane < w.qlen;
                const uint32_t a0 = act ? w.q_ka[lane] : 0u, b0 = act ? w.q_kb[lane] : 0u, u0 = act ? w.q_u[lane] : 0u;
                fz_resolve<REAL>(t, hv, w, act, a0, b0, u0);
                w.qlen = 0;
            }
            double sr = w.er, si = w.ei;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, d);
                si += __shfl_xor_sync(0xffffffffu, si, d);
            }
            if (lane == 0) {
                const double2 a = amps[row_start + r];
                const double den = a.x * a.x + a.y * a.y;
                eloc[r] = make_double2((sr * a.x + si * a.y) / den, (si * a.x - sr * a.y) / den);
            }
        }
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_local_energy_sample_aware(const anqs_tables_t *h, const int64_t *d_samples, const double *d_amps,
                                   int64_t n_total, int64_t row_start, int64_t row_len, const void *d_table,
                                   int64_t capacity, int alpha_num, int beta_num, double *d_eloc, void *stream) {
    return anqs_local_energy_sample_aware_variant(h, d_samples, d_amps, n_total, row_start, row_len, d_table, capacity, alpha_num, beta_num,
                                                  d_eloc, 0, stream);
}

int anqs_local_energy_sample_aware_variant(const anqs_tables_t *h, const int64_t *d_samples, const double *d_amps,
                                           int64_t n_total, int64_t row_start, int64_t row_len, const void *d_table,
                                           int64_t capacity, int alpha_num, int beta_num, double *d_eloc, int variant, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(variant >= 0 && variant <= 2, "variant must be 0 (automatic), 1 (warp-per-sample kernel) or 2 (bit-sliced kernel)");
    ANQS_REQUIRE(row_start >= 0 && row_len >= 0 && row_start + row_len <= n_total, "row window out of range");
    if (row_len == 0) return 0;
    ANQS_REQUIRE(d_samples && d_amps && d_table && d_eloc, "null pointer");
    ANQS_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 1024");
    const Tables *t = (const Tables *)h;
    const size_t smem = (size_t)FZ_QUEUE_BYTES + (size_t)t->tile_bytes_max;
    HashView hv = make_hash_view(d_table, capacity);
    const int64_t ngroups = (row_len + FZ_WARPS - 1) / FZ_WARPS;
    const int grid = (int)std::min<int64_t>(ngroups, sm_count_of_current_device());
    cudaStream_t s = (cudaStream_t)stream;
    {   // bit-sliced kernel (k1_fused_bs.cu) whenever the table allows it
        const int rc = fused_bs_try_launch(t, hv, d_samples, d_amps, row_start, row_len, alpha_num, beta_num, d_eloc, variant, s);
        ANQS_REQUIRE(rc >= 0, "cudaFuncSetAttribute failed for the bit-sliced kernel");
        if (rc == 1) {
            ANQS_LAUNCH_CHECK();
            return 0;
        }
    }
    if (t->weights_real) {
        auto kern = fused_eloc_kernel<true>;
        ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, FZ_THREADS, smem, s>>>(*t, hv, d_samples, (const double2 *)d_amps, row_start, row_len, alpha_num,
                                            beta_num, (double2 *)d_eloc);
    } else {
        auto kern = fused_eloc_kernel<false>;
        ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, FZ_THREADS, smem, s>>>(*t, hv, d_samples, (const double2 *)d_amps, row_start, row_len, alpha_num,
                                            beta_num, (double2 *)d_eloc);
    }
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
