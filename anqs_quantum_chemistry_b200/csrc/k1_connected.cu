// Kernel family 1: connected-configuration generation against the Jordan-Wigner Hamiltonian tables.
//
//   k1_filter_kernel   x' = x ^ xy[u] for every unique XY mask, alpha/beta electron-count filter
//                      (reference PO:527-567), result = per-sample count + ballot bitmap
//   k1_emit_kernel     bitmap -> ordered list (dest, x', xy_ptr, H_{x,x'})   (PO:527-567 + PO:256-324)
//   matrix_elements    H_{x,x'} for an arbitrary (x', xy_ptr) list              (PO:256-324)
//   (the fused sample-aware local-energy kernel lives in k1_fused.cu)
//   accumulate_rows    E[dest] += H * psi(src) over a CSR list                  (PO:453-478)
//
// Data layout: the XY masks are de-interleaved once at table-build time into (even bits, odd bits) =
// (alpha, beta) 32-bit words.  The reference's filter popcount(x' & 0x5555..) == N_alpha becomes
// popcount(xa ^ ma) == N_alpha: one XOR + one 32-bit POPC per spin sector instead of two 64-bit
// popcounts.  A warp owns one sample (or SPW samples) and sweeps the mask table, 32 masks per step, from
// shared memory; the table is staged by 1-D bulk TMA copies (cp.async.bulk + mbarrier), resident when it
// fits (<= 200 KB) and double-buffered 64 KB tiles otherwise.  Ballots give ordered compaction for free,
// so the emitted list is lexicographic in (dest, xy_ptr) like the reference's.
#include <algorithm>

#include "common.cuh"
#include "matrix_elements.cuh"

namespace anqs {

constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int TILE_STREAM = 8192;         // masks per streamed tile (64 KB)
constexpr int TILE_RESIDENT_MAX = 24576;  // masks kept resident (192 KB); multiple of 1024

// ---- mask-table staging ------------------------------------------------------------------------
struct TileStager {
    uint2 *buf;          // [nbuf][tile_masks]
    uint64_t *bars;      // [2]
    uint32_t parity[2];
    const uint2 *gmem;
    int64_t U_pad;
    int tile_masks, ntiles;
    bool resident_loaded;

    __device__ void init(uint2 *b, uint64_t *br, const uint2 *g, int64_t upad, int tm, int nt) {
        buf = b; bars = br; gmem = g; U_pad = upad; tile_masks = tm; ntiles = nt;
        parity[0] = parity[1] = 0;
        resident_loaded = false;
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    __device__ int tile_len(int t) const {
        int64_t rem = U_pad - (int64_t)t * tile_masks;
        return (int)(rem < tile_masks ? rem : tile_masks);
    }
    __device__ void issue(int t) {
        if (threadIdx.x == 0) {
            int b = t & 1;
            uint32_t bytes = (uint32_t)tile_len(t) * 8u;
            mbar_arrive_expect_tx(&bars[b], bytes);
            // one bulk copy may move at most 2^20-16 bytes; tiles are <= 200 KB
            bulk_copy_g2s(buf + (size_t)b * tile_masks, gmem + (size_t)t * tile_masks, bytes, &bars[b]);
        }
    }
    __device__ const uint2 *wait(int t) {
        int b = t & 1;
        mbar_wait(&bars[b], parity[b]);
        parity[b] ^= 1u;
        return buf + (size_t)b * tile_masks;
    }
};

// ---- kernel 1a: filter ---------------------------------------------------------------------------
template <int SPW, bool CHECK>
__device__ __forceinline__ void filter_steps(const uint2 *tile, int it_begin, int it_end, int64_t base_idx, int64_t U,
                                             const uint32_t (&xa)[SPW], const uint32_t (&xb)[SPW], int alpha, int beta,
                                             int (&cnt)[SPW], uint32_t (&keep)[SPW], uint32_t *bitmap, int64_t row_words,
                                             int64_t s0, int64_t n) {
    const int lane = lane_id();
#pragma unroll 2
    for (int it = it_begin; it < it_end; ++it) {
        int i = it * 32 + lane;
        uint2 m = tile[i];
        bool valid = CHECK ? (base_idx + i < U) : true;
#pragma unroll
        for (int k = 0; k < SPW; ++k) {
            bool p = valid && (__popc(xa[k] ^ m.x) == alpha) && (__popc(xb[k] ^ m.y) == beta);
            unsigned b = __ballot_sync(0xffffffffu, p);
            cnt[k] += __popc(b);
            if (lane == (it & 31)) keep[k] = b;
        }
        if ((it & 31) == 31 && bitmap) {
#pragma unroll
            for (int k = 0; k < SPW; ++k)
                if (s0 + k < n) bitmap[(s0 + k) * row_words + (base_idx >> 5) + (it & ~31) + lane] = keep[k];
        }
    }
}

template <int SPW>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_filter_kernel(const uint2 *__restrict__ mab, int64_t U, int64_t U_pad, const int64_t *__restrict__ samples, int64_t n,
                 int alpha, int beta, int64_t *__restrict__ counts, uint32_t *__restrict__ bitmap, int64_t row_words,
                 int tile_masks, int ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[2];
    TileStager st;
    st.init(reinterpret_cast<uint2 *>(smem_raw), bars, mab, U_pad, tile_masks, ntiles);

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int64_t per_group = (int64_t)K1_WARPS * SPW;
    const int64_t ngroups = (n + per_group - 1) / per_group;
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
        const int64_t s0 = group * per_group + (int64_t)warp * SPW;
        uint32_t xa[SPW], xb[SPW], keep[SPW];
        int cnt[SPW];
#pragma unroll
        for (int k = 0; k < SPW; ++k) {
            uint64_t x = (s0 + k < n) ? (uint64_t)samples[s0 + k] : 0ull;
            xa[k] = compress_even_bits(x);
            xb[k] = compress_even_bits(x >> 1);
            cnt[k] = 0;
            keep[k] = 0;
        }
        if (ntiles > 1) st.issue(0);
        for (int t = 0; t < ntiles; ++t) {
            const uint2 *tile;
            if (ntiles == 1) {
                if (!st.resident_loaded) {
                    st.issue(0);
                    tile = st.wait(0);
                    st.resident_loaded = true;
                } else {
                    tile = st.buf;
                }
            } else {
                if (t + 1 < ntiles) st.issue(t + 1);
                tile = st.wait(t);
            }
            const int tl = st.tile_len(t);
            const int64_t base_idx = (int64_t)t * tile_masks;
            int64_t valid_masks = U - base_idx;
            if (valid_masks < 0) valid_masks = 0;
            if (valid_masks > tl) valid_masks = tl;
            const int full_its = (int)(valid_masks >> 5), all_its = tl >> 5;
            filter_steps<SPW, false>(tile, 0, full_its, base_idx, U, xa, xb, alpha, beta, cnt, keep, bitmap, row_words, s0, n);
            filter_steps<SPW, true>(tile, full_its, all_its, base_idx, U, xa, xb, alpha, beta, cnt, keep, bitmap, row_words, s0, n);
            if (ntiles > 1) __syncthreads();  // everyone is done with this buffer before it is refilled
        }
        if (counts && lane == 0) {
#pragma unroll
            for (int k = 0; k < SPW; ++k)
                if (s0 + k < n) counts[s0 + k] = cnt[k];
        }
    }
}

// ---- kernel 1b: emit --------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;
constexpr int EMIT_QCAP = 1024 + 32;

template <bool REAL, int HC>
__device__ __forceinline__ void emit_batch(const Tables &t, uint64_t x, int s, bool active, uint32_t u, int64_t r,
                                           int32_t *dest, int64_t *xprime, int32_t *xy_ptr, double *H) {
    uint64_t xp = 0;
    int2 g = make_int2(0, 0);
    if (active) {
        xp = x ^ __ldg(t.xy + u);
        if (HC) g = __ldg(t.grp + u);
    }
    double hr = 0.0, hi = 0.0;
    if (HC) warp_matrix_elements<REAL>(t, active, g, deinterleave(xp), hr, hi);
    if (active) {
        if (dest) dest[r] = s;
        xprime[r] = (int64_t)xp;
        if (xy_ptr) xy_ptr[r] = (int32_t)u;
        if (HC == 1) H[r] = hr;
        if (HC == 2) reinterpret_cast<double2 *>(H)[r] = make_double2(hr, hi);
    }
}

template <bool REAL, int HC>
__global__ void __launch_bounds__(EMIT_THREADS)
k1_emit_kernel(Tables t, const int64_t *__restrict__ samples, int64_t n, const uint32_t *__restrict__ bitmap,
               const int64_t *__restrict__ offsets, int32_t *__restrict__ dest, int64_t *__restrict__ xprime,
               int32_t *__restrict__ xy_ptr, double *__restrict__ H) {
    __shared__ uint32_t queue_all[EMIT_WARPS][EMIT_QCAP];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint32_t *q = queue_all[warp];
    const int words = (int)((t.U + 31) >> 5);
    for (int64_t s = (int64_t)blockIdx.x * EMIT_WARPS + warp; s < n; s += (int64_t)gridDim.x * EMIT_WARPS) {
        const uint64_t x = (uint64_t)samples[s];
        int64_t out = offsets[s];
        int qlen = 0;
        const uint32_t *row = bitmap + s * t.row_words;
        for (int j = 0; j < words; j += 32) {
            uint32_t w = (j + lane < words) ? __ldg(row + j + lane) : 0u;
            int c = __popc(w);
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            int total = __shfl_sync(0xffffffffu, inc, 31);
            int p = qlen + inc - c;
            while (w) {
                int bit = __ffs(w) - 1;
                w &= w - 1;
                q[p++] = (uint32_t)(((j + lane) << 5) + bit);
            }
            __syncwarp();
            qlen += total;
            int done = 0;
            while (qlen - done >= 32) {
                emit_batch<REAL, HC>(t, x, (int)s, true, q[done + lane], out + done + lane, dest, xprime, xy_ptr, H);
                done += 32;
            }
            if (done > 0) {
                int rem = qlen - done;
                uint32_t v = lane < rem ? q[done + lane] : 0u;
                __syncwarp();
                if (lane < rem) q[lane] = v;
                __syncwarp();
                out += done;
                qlen = rem;
            }
        }
        if (qlen > 0) emit_batch<REAL, HC>(t, x, (int)s, lane < qlen, lane < qlen ? q[lane] : 0u, out + lane, dest, xprime, xy_ptr, H);
        __syncwarp();
    }
}

// ---- PO:256-324 on an arbitrary list -----------------------------------------------------------------
template <bool REAL>
__global__ void __launch_bounds__(256)
matrix_elements_kernel(Tables t, const int64_t *__restrict__ xprime, const int64_t *__restrict__ xy_ptr, int64_t m,
                       double2 *__restrict__ H) {
    int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; base < m; base += stride) {
        int64_t i = base + lane_id();
        bool active = i < m;
        uint64_t xp = 0;
        int2 g = make_int2(0, 0);
        if (active) {
            xp = deinterleave((uint64_t)xprime[i]);
            g = __ldg(t.grp + xy_ptr[i]);
        }
        double hr, hi;
        warp_matrix_elements<REAL>(t, active, g, xp, hr, hi);
        if (active) H[i] = make_double2(hr, hi);
    }
}

// ---- E[dest] += H * psi(src) over CSR rows ---------------------------------------------------------------
template <int HC>
__global__ void __launch_bounds__(256)
accumulate_rows_kernel(const int64_t *__restrict__ offsets, int64_t n, const int64_t *__restrict__ src_ptr,
                       const double *__restrict__ H, const double2 *__restrict__ src_amps,
                       const double2 *__restrict__ amps_dest, double2 *__restrict__ eloc, int accumulate) {
    const int lane = lane_id();
    int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = wid; i < n; i += nw) {
        double sr = 0.0, si = 0.0;
        for (int64_t r = offsets[i] + lane; r < offsets[i + 1]; r += 32) {
            int64_t p = src_ptr[r];
            if (p < 0) continue;
            double2 a = src_amps[p];
            if (HC == 1) {
                double h = H[r];
                sr += h * a.x;
                si += h * a.y;
            } else {
                double2 h = reinterpret_cast<const double2 *>(H)[r];
                sr += h.x * a.x - h.y * a.y;
                si += h.x * a.y + h.y * a.x;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, d);
            si += __shfl_xor_sync(0xffffffffu, si, d);
        }
        if (lane == 0) {
            if (accumulate) {
                sr += eloc[i].x;
                si += eloc[i].y;
            }
            if (amps_dest) {
                double2 a = amps_dest[i];
                double den = a.x * a.x + a.y * a.y;
                eloc[i] = make_double2((sr * a.x + si * a.y) / den, (si * a.x - sr * a.y) / den);
            } else {
                eloc[i] = make_double2(sr, si);
            }
        }
    }
}

static void pick_tiling(const Tables *t, int *tile_masks, int *ntiles, size_t *smem) {
    if (t->U_pad <= TILE_RESIDENT_MAX) {
        *tile_masks = (int)t->U_pad;
        *ntiles = 1;
        *smem = (size_t)t->U_pad * sizeof(uint2);
    } else {
        *tile_masks = TILE_STREAM;
        *ntiles = (int)((t->U_pad + TILE_STREAM - 1) / TILE_STREAM);
        *smem = (size_t)2 * TILE_STREAM * sizeof(uint2);
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_k1_filter(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                   int64_t *d_counts, uint32_t *d_bitmap, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_samples, "null samples");
    ANQS_REQUIRE(d_counts || d_bitmap, "nothing to compute: both outputs are NULL");
    const Tables *t = (const Tables *)h;
    constexpr int SPW = 4;
    int tile_masks, ntiles;
    size_t smem;
    pick_tiling(t, &tile_masks, &ntiles, &smem);
    auto kern = k1_filter_kernel<SPW>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ngroups = (n + K1_WARPS * SPW - 1) / (K1_WARPS * SPW);
    int grid = (int)std::min<int64_t>(ngroups, sm_count_of_current_device());
    kern<<<grid, K1_THREADS, smem, (cudaStream_t)stream>>>(t->mab, t->U, t->U_pad, d_samples, n, alpha_num, beta_num,
                                                           d_counts, d_bitmap, t->row_words, tile_masks, ntiles);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_k1_emit(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, const uint32_t *d_bitmap,
                 const int64_t *d_offsets, int32_t *d_dest, int64_t *d_xprime, int32_t *d_xy_ptr, double *d_H,
                 int h_components, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(n < ((int64_t)1 << 31), "chunk too large for int32 dest; split the batch");
    ANQS_REQUIRE(d_samples && d_bitmap && d_offsets && d_xprime, "null pointer");
    const Tables *t = (const Tables *)h;
    int hc = d_H ? h_components : 0;
    ANQS_REQUIRE(hc == 0 || hc == 1 || hc == 2, "h_components must be 1 (real) or 2 (complex)");
    ANQS_REQUIRE(!(hc == 1 && !t->weights_real), "real matrix elements requested but the Hamiltonian weights are complex");
    int grid = (int)std::min<int64_t>((n + EMIT_WARPS - 1) / EMIT_WARPS, (int64_t)sm_count_of_current_device() * 8);
    cudaStream_t s = (cudaStream_t)stream;
#define ANQS_EMIT(REAL, HC) \
    k1_emit_kernel<REAL, HC><<<grid, EMIT_THREADS, 0, s>>>(*t, d_samples, n, d_bitmap, d_offsets, d_dest, d_xprime, d_xy_ptr, d_H)
    if (hc == 0) ANQS_EMIT(true, 0);
    else if (t->weights_real && hc == 1) ANQS_EMIT(true, 1);
    else if (t->weights_real && hc == 2) ANQS_EMIT(true, 2);
    else ANQS_EMIT(false, 2);
#undef ANQS_EMIT
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_matrix_elements(const anqs_tables_t *h, const int64_t *d_xprime, const int64_t *d_xy_ptr, int64_t m,
                         double *d_H, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(m >= 0, "negative element count");
    if (m == 0) return 0;
    ANQS_REQUIRE(d_xprime && d_xy_ptr && d_H, "null pointer");
    const Tables *t = (const Tables *)h;
    int grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (t->weights_real)
        matrix_elements_kernel<true><<<grid, 256, 0, s>>>(*t, d_xprime, d_xy_ptr, m, (double2 *)d_H);
    else
        matrix_elements_kernel<false><<<grid, 256, 0, s>>>(*t, d_xprime, d_xy_ptr, m, (double2 *)d_H);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_accumulate_rows(const int64_t *d_offsets, int64_t n, const int64_t *d_src_ptr, const double *d_H,
                         int h_components, const double *d_src_amps, const double *d_amps_dest, double *d_eloc,
                         int accumulate, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative row count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_offsets && d_src_ptr && d_H && d_src_amps && d_eloc, "null pointer");
    ANQS_REQUIRE(h_components == 1 || h_components == 2, "h_components must be 1 or 2");
    int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)sm_count_of_current_device() * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (h_components == 1)
        accumulate_rows_kernel<1><<<grid, 256, 0, s>>>(d_offsets, n, d_src_ptr, d_H, (const double2 *)d_src_amps,
                                                       (const double2 *)d_amps_dest, (double2 *)d_eloc, accumulate);
    else
        accumulate_rows_kernel<2><<<grid, 256, 0, s>>>(d_offsets, n, d_src_ptr, d_H, (const double2 *)d_src_amps,
                                                       (const double2 *)d_amps_dest, (double2 *)d_eloc, accumulate);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
