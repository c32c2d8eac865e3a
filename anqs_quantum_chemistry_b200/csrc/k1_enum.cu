// Kernel family 1, tiled variant: the ordered connected list (dest, x', xy_ptr, H_{x,x'}) of PO:527-567 + PO:256-324 at
// HBM-write speed.  Same results and the same output order as k1_filter_kernel + k1_emit_kernel (k1_connected.cu); what
// changes is where the bytes come from:
//
//   enum_filter_kernel   alpha/beta electron-count filter through the PRODUCT layout of the masks (rows of masks that share
//                        their alpha part: ~A alpha tests + ~0.4 U beta tests per sample instead of 2 U tests).  The
//                        passing masks set bits of a per-warp bitmap row in shared memory (ATOMS.OR), which is then written
//                        out in one piece together with the sample's count and the rank of every enumeration tile's first
//                        connection inside the sample (tile_prefix).
//   enum_emit_kernel     every CTA keeps ONE enumeration tile resident in shared memory - a contiguous range of masks with
//                        their XY words, YZ-group descriptors and the 16-byte term records of those groups, staged by bulk
//                        TMA copies - and streams samples past it: the tile's slice of the bitmap row gives the passing masks
//                        in order, tile_prefix their position in the output, and the matrix-element sums read their term
//                        records from shared memory instead of L2 (the L2 path of the untiled kernel moves ~4x the bytes
//                        that go to HBM and caps it at 13 % of the HBM roofline).  Samples are handed out per warp from one
//                        atomic counter per tile; a CTA that runs out of samples for its tile moves on to the next tile
//                        that still has some, so heavy tiles (the one holding the diagonal group) get helped.
//
// Algorithmic HBM traffic: 20 B written per emitted connection (x' 8, H 8, dest 4; +4 with xy_ptr, +8 with complex H), the
// bitmap row (U/8 B) written once and read once per sample, 8 B per sample in, 4 B per (sample, tile) of ranks.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "matrix_elements.cuh"

namespace anqs {

constexpr int EN_THREADS = 1024;
constexpr int EN_WARPS = EN_THREADS / 32;
constexpr int EN_SMEM_MAX = 227 * 1024;
constexpr uint32_t EN_BULK_CHUNK = 64 * 1024;

__device__ __forceinline__ void stage_blob(unsigned char *dst, const uint8_t *src, uint32_t bytes, uint64_t *bar) {
    mbar_arrive_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += EN_BULK_CHUNK) bulk_copy_g2s(dst + off, src + off, min(EN_BULK_CHUNK, bytes - off), bar);
}

// ---- pass 1: product-layout filter -> bitmap rows, counts, per-tile ranks ------------------------------------------------
__device__ __forceinline__ void ef_set(uint32_t *bm, uint32_t u) { atomicOr(bm + (u >> 5), 1u << (u & 31u)); }

__device__ __forceinline__ void ef_process_tile(uint32_t *bm, uint32_t xa, uint32_t xb, int alpha, int beta, const ProdTile &tile,
                                                const unsigned char *smem_tile) {
    const int lane = lane_id();
    const uint4 *rows = reinterpret_cast<const uint4 *>(smem_tile);
    const uint2 *mems = reinterpret_cast<const uint2 *>(smem_tile + (size_t)(tile.n_multi + tile.n_single) * sizeof(RowRec));
    for (uint32_t r0 = 0; r0 < tile.n_multi; r0 += 32) {
        const uint32_t r = r0 + lane;
        bool pass_a = false;
        if (r < tile.n_multi) pass_a = __popc(xa ^ rows[r].x) == alpha;
        unsigned todo = __ballot_sync(0xffffffffu, pass_a);
        while (todo) {
            const uint32_t rr = r0 + (__ffs(todo) - 1);
            todo &= todo - 1;
            const uint4 rec = rows[rr];  // uniform address: one broadcast load
            const uint2 *m = mems + rec.z;
            const uint32_t len = rec.w;
            for (uint32_t j0 = 0; j0 < len; j0 += 64) {
                const uint32_t j1 = j0 + lane, j2 = j0 + 32 + lane;
                uint2 m1 = make_uint2(0, 0), m2 = make_uint2(0, 0);
                if (j1 < len) m1 = m[j1];
                if (j2 < len) m2 = m[j2];
                if (j1 < len && __popc(xb ^ m1.x) == beta) ef_set(bm, m1.y);
                if (j2 < len && __popc(xb ^ m2.x) == beta) ef_set(bm, m2.y);
            }
        }
    }
    const uint4 *srows = rows + tile.n_multi;
    for (uint32_t r0 = 0; r0 < tile.n_single; r0 += 64) {
        const uint32_t r1 = r0 + lane, r2 = r0 + 32 + lane;
        uint4 c1 = make_uint4(0, 0, 0, 0), c2 = make_uint4(0, 0, 0, 0);
        if (r1 < tile.n_single) c1 = srows[r1];
        if (r2 < tile.n_single) c2 = srows[r2];
        if (r1 < tile.n_single && __popc(xa ^ c1.x) == alpha && __popc(xb ^ c1.z) == beta) ef_set(bm, c1.w);
        if (r2 < tile.n_single && __popc(xa ^ c2.x) == alpha && __popc(xb ^ c2.z) == beta) ef_set(bm, c2.w);
    }
}

__global__ void __launch_bounds__(EN_THREADS, 1)
enum_filter_kernel(Tables t, const int64_t *__restrict__ samples, int64_t n, int alpha, int beta, int64_t *__restrict__ counts,
                   uint32_t *__restrict__ bitmap, int32_t *__restrict__ tile_prefix, uint32_t bitmap_smem_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ ProdTile s_tile;
    const int warp = threadIdx.x >> 5, lane = lane_id(), nwarps = blockDim.x >> 5;
    const int row_words = (int)t.row_words;
    uint32_t *bm = reinterpret_cast<uint32_t *>(smem_raw) + (size_t)warp * row_words;
    unsigned char *tile_buf = smem_raw + bitmap_smem_bytes;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    const bool resident = t.n_tiles == 1;
    bool loaded = false;
    const int64_t ngroups = (n + nwarps - 1) / nwarps;
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
        const int64_t r = group * nwarps + warp;
        const bool have = r < n;
        const uint64_t x = have ? (uint64_t)samples[r] : 0ull;
        const uint32_t xa = compress_even_bits(x), xb = compress_even_bits(x >> 1);
        for (int j = lane; j < row_words; j += 32) bm[j] = 0u;
        __syncwarp();
        for (int ti = 0; ti < t.n_tiles; ++ti) {
            if (!resident || !loaded) {
                __syncthreads();  // everyone is done with the previous contents of tile_buf / s_tile
                if (threadIdx.x == 0) {
                    const ProdTile pt = t.prod_tiles[ti];
                    s_tile = pt;
                    stage_blob(tile_buf, t.prod_blob_u + pt.blob_off, pt.blob_bytes, &bar);
                }
                __syncthreads();
                mbar_wait(&bar, parity);
                parity ^= 1u;
                loaded = true;
            }
            const ProdTile tile = s_tile;
            if (have) ef_process_tile(bm, xa, xb, alpha, beta, tile, tile_buf);
        }
        __syncwarp();
        if (have) {
            int total = 0;
            if (tile_prefix) {
                for (int e = 0; e < t.n_enum_tiles; ++e) {
                    const int w0 = (int)__ldg(&t.enum_tiles[e].word0), nw = (int)__ldg(&t.enum_tiles[e].n_words);
                    int c = 0;
                    for (int j = w0 + lane; j < w0 + nw; j += 32) c += __popc(bm[j]);
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == 0) tile_prefix[r * t.n_enum_tiles + e] = total;
                    total += c;
                }
            } else {
                int c = 0;
                for (int j = lane; j < row_words; j += 32) c += __popc(bm[j]);
                total = __reduce_add_sync(0xffffffffu, c);
            }
            if (counts && lane == 0) counts[r] = total;
            if (bitmap) {
                uint32_t *row = bitmap + r * t.row_words;
                for (int j = lane; j < row_words; j += 32) row[j] = bm[j];
            }
        }
        __syncwarp();
    }
}

// ---- pass 1, bit-sliced: 32 samples per lane-operation -------------------------------------------------------------------
// For a sample inside the (N_alpha, N_beta) sector, x' = x ^ xy keeps the electron counts iff exactly half of the positions of
// the alpha part of the mask, and half of those of its beta part, are occupied in x.  With the samples of a group of 32
// BIT-SLICED (slice i = bit i of the 32 samples, one ballot each), that test is a boolean function of <= 8 slices evaluated
// for 32 samples at once: one lane = one mask, ~20 LOP3 and no POPC per 32 (sample, mask) pairs.  A 32 x 32 bit transpose
// across the warp turns the 32 result words (mask-major) into the bitmap words of the 32 samples.  Samples outside the
// sector (never produced by the masked samplers, but legal input) get their rows recomputed with the plain popcount test.
constexpr int BS_THREADS = 512;
constexpr int BS_WARPS = BS_THREADS / 32;

__device__ __forceinline__ uint32_t exactly_two(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return ~(a ^ b ^ c ^ d) & ~(a & b & c & d) & (a | b | c | d);
}

// lane l holds row l of a 32 x 32 bit matrix; afterwards lane l holds column l
__device__ __forceinline__ uint32_t transpose32(uint32_t w) {
    const int lane = lane_id();
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        const uint32_t mk = k == 16 ? 0x0000FFFFu : k == 8 ? 0x00FF00FFu : k == 4 ? 0x0F0F0F0Fu : k == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t p = __shfl_xor_sync(0xffffffffu, w, k);
        w = (lane & k) ? ((w & ~mk) | ((p >> k) & mk)) : ((w & mk) | ((p & mk) << k));
    }
    return w;
}

__global__ void __launch_bounds__(BS_THREADS, 2)
enum_filter_bitsliced_kernel(Tables t, const int64_t *__restrict__ samples, int64_t n, int alpha, int beta,
                             int64_t *__restrict__ counts, uint32_t *__restrict__ bitmap, int32_t *__restrict__ tile_prefix,
                             int chunk_words) {
    // The bitmap rows of the 32 samples are built chunk_words words at a time in shared memory (one chunk when the whole row
    // fits, as for every BASELINE table; tables with very many masks take several).
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t X[66];
    __shared__ uint64_t s_x[32];
    __shared__ uint32_t s_bad, s_valid;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int row_words = (int)t.row_words, stride = chunk_words | 1;  // odd stride: the 32 rows of a column hit 32 banks
    const int n_tiles = t.n_enum_tiles;
    uint32_t *rows = reinterpret_cast<uint32_t *>(smem_raw);
    const int64_t ngroups = (n + 31) >> 5;
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
        {   // slices: every warp reads the 32 samples and builds 64 / BS_WARPS of the slices
            const int64_t r = group * 32 + lane;
            const bool valid = r < n;
            const uint64_t x = valid ? (uint64_t)samples[r] : 0ull;
            constexpr int PER = 64 / BS_WARPS;
            uint32_t mine = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const uint32_t b = __ballot_sync(0xffffffffu, (x >> (warp * PER + i)) & 1ull);
                if (lane == i) mine = b;
            }
            if (lane < PER) X[warp * PER + lane] = mine;
            if (warp == 0) {
                const bool insec = __popcll(x & 0x5555555555555555ULL) == alpha && __popcll(x & 0xAAAAAAAAAAAAAAAAULL) == beta;
                s_x[lane] = x;
                const uint32_t bad = __ballot_sync(0xffffffffu, valid && !insec), val = __ballot_sync(0xffffffffu, valid);
                if (lane == 0) {
                    X[64] = 0u;
                    X[65] = 0xffffffffu;
                    s_bad = bad;
                    s_valid = val;
                }
            }
        }
        __syncthreads();
        const uint32_t bad = s_bad, valid = s_valid;
        int e_first = 0;  // first enumeration tile that reaches into the current chunk
        for (int c0 = 0; c0 < row_words; c0 += chunk_words) {
            const int c1 = min(c0 + chunk_words, row_words);
            for (int j0 = c0 + warp; j0 < c1; j0 += BS_WARPS * 4) {
                uint2 pos[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = j0 + k * BS_WARPS;
                    pos[k] = j < c1 ? __ldg(t.bs_pos + (size_t)j * 32 + lane) : make_uint2(0x40404040u, 0x40404040u);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = j0 + k * BS_WARPS;
                    if (j >= c1) break;
                    const uint32_t fa = exactly_two(X[pos[k].x & 0xff], X[(pos[k].x >> 8) & 0xff], X[(pos[k].x >> 16) & 0xff], X[pos[k].x >> 24]);
                    const uint32_t fb = exactly_two(X[pos[k].y & 0xff], X[(pos[k].y >> 8) & 0xff], X[(pos[k].y >> 16) & 0xff], X[pos[k].y >> 24]);
                    rows[lane * stride + (j - c0)] = transpose32(fa & fb);
                }
            }
            __syncthreads();
            if (bad) {  // plain popcount test for the samples outside the sector (whole warp per sample)
                for (int sidx = warp; sidx < 32; sidx += BS_WARPS) {
                    if (!((bad >> sidx) & 1u)) continue;
                    const uint64_t x = s_x[sidx];
                    const uint32_t xa = compress_even_bits(x), xb = compress_even_bits(x >> 1);
                    for (int j = c0; j < c1; ++j) {
                        const uint2 m = __ldg(t.mab + (size_t)j * 32 + lane);
                        const bool p = (int64_t)j * 32 + lane < t.U && __popc(xa ^ m.x) == alpha && __popc(xb ^ m.y) == beta;
                        const uint32_t b = __ballot_sync(0xffffffffu, p);
                        if (lane == 0) rows[sidx * stride + (j - c0)] = b;
                    }
                }
                __syncthreads();
            }
            while (e_first < n_tiles && (int)(__ldg(&t.enum_tiles[e_first].word0) + __ldg(&t.enum_tiles[e_first].n_words)) <= c0) ++e_first;
            for (int sidx = warp; sidx < 32; sidx += BS_WARPS) {
                if (!((valid >> sidx) & 1u)) continue;
                const int64_t r = group * 32 + sidx;
                const uint32_t *bm = rows + sidx * stride - c0;  // bm[j] = word j of the row, c0 <= j < c1
                uint32_t *row = bitmap + r * t.row_words;
                // one pass over the row's words: written out to the bitmap and counted per enumeration tile that overlaps this
                // chunk (raw counts for now, scanned below); the tiles cover every word of a row
                for (int e = e_first; e < n_tiles; ++e) {
                    const int w0 = (int)__ldg(&t.enum_tiles[e].word0), w1 = w0 + (int)__ldg(&t.enum_tiles[e].n_words);
                    if (w0 >= c1) break;
                    int c = 0;
                    for (int j = max(w0, c0) + lane; j < min(w1, c1); j += 32) {
                        const uint32_t v = bm[j];
                        row[j] = v;
                        c += __popc(v);
                    }
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == 0) {
                        int32_t *slot = tile_prefix + r * n_tiles + e;
                        *slot = w0 >= c0 ? c : *slot + c;  // a tile that started in an earlier chunk already has a partial count
                    }
                }
            }
            __syncthreads();
        }
        // per-tile counts -> rank of every tile's first connection inside the sample (exclusive scan), total -> counts
        for (int sidx = warp; sidx < 32; sidx += BS_WARPS) {
            if (!((valid >> sidx) & 1u)) continue;
            const int64_t r = group * 32 + sidx;
            int32_t *tp = tile_prefix + r * n_tiles;
            int carry = 0;
            for (int e0 = 0; e0 < n_tiles; e0 += 32) {
                const int e = e0 + lane;
                const int v = e < n_tiles ? tp[e] : 0;
                int inc = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += o;
                }
                if (e < n_tiles) tp[e] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) counts[r] = carry;
        }
        __syncthreads();
    }
}

// ---- pass 2: tile-resident ordered emit ------------------------------------------------------------------------------
// A warp owns one (sample, tile) unit at a time.  The tile's slice of the sample's bitmap row is expanded ENUM_STEP_WORDS words
// at a time into a queue of tile-local mask indices (ascending = output order); every 32 queued indices are one batch, one
// connection per lane: one LDS.128 {xy, mult, table offset} + one LDS.64 table entry, ~25 integer
// instructions, three coalesced streaming stores.  Generic connections (mult = 0, ~1 %) are appended to the warp's deferred
// list (global workspace, L2) and evaluated 32 at a time from the occupation blocks of the tile; samples outside the sector
// take et_unit_slow.
constexpr int ET_STEP_WORDS = ENUM_STEP_WORDS;
constexpr int ET_QCAP = ENUM_QCAP;
constexpr int ET_QUEUE_BYTES = ENUM_QUEUE_BYTES;
static_assert(EN_WARPS == ENUM_EMIT_WARPS, "queue budget");

__device__ __forceinline__ double flip_hi(double w, uint32_t sign) {
    return __hiloint2double(__double2hiint(w) ^ (int)sign, __double2loint(w));
}
// shared-memory loads by 32-bit shared address (no generic-to-shared conversion in the inner loop); the tile is read-only
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

struct EtOut {  // per-lane output cursors of the unit being emitted
    int32_t *dest;
    long long *xprime;
    int32_t *xy_ptr;
    double *H;
};

__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}

// Inclusive warp scan of the per-lane popcounts of one step of bitmap words and expansion of the set bits into the queue
// (ascending; q_s = shared address of the warp's queue).  Returns the number of indices appended.
__device__ __forceinline__ int et_expand(uint32_t q_s, int qlen, uint32_t w, int j, int &lane_start) {
    const int lane = lane_id();
    const int c = __popc(w);
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    lane_start = qlen + inc - c;  // queue position of the lane's first index
    if (total == 0) return 0;
    // The lane's indices go to q[qlen + inc - c .. qlen + inc), two per iteration: the lowest remaining bit from the front, the
    // highest from the back (for an odd count the middle one is written twice, to the same slot).  The trip count is the
    // warp's maximum, so the loop branch is uniform and the body predicated: no reconvergence bookkeeping per iteration.
    uint32_t qa_lo = q_s + 2u * (uint32_t)(qlen + inc - c), qa_hi = q_s + 2u * (uint32_t)(qlen + inc);
    const uint32_t base = (uint32_t)(j + lane) << 5;
    const int trips = (__reduce_max_sync(0xffffffffu, c) + 1) >> 1;
    for (int it = 0; it < trips; ++it) {
        if (w) {
            const uint32_t hb = 31u - (uint32_t)__clz((int)w), lb = (uint32_t)__ffs((int)w) - 1u;
            qa_hi -= 2u;
            sts_u16(qa_hi, base + hb);
            sts_u16(qa_lo, base + lb);
            qa_lo += 2u;
            w &= ~(1u << hb);
            w &= w - 1u;
        }
    }
    __syncwarp();
    return total;
}

// Matrix elements of up to 32 deferred connections of the resident tile, one per lane (whole warp).
// ent.x = x', ent.y = (output position << 16) | tile-local mask index.
template <bool REAL, int HC>
__device__ __noinline__ void et_eval_deferred(const Tables *td, uint32_t tile_s, uint32_t u0, ulonglong2 ent, bool active,
                                              double *H) {
    const Tables &t = *td;
    const int lane = lane_id(), n = t.qubit_num;
    const uint32_t k = (uint32_t)ent.y & 0xffffu;
    const uint32_t plo = (uint32_t)ent.x, phi = (uint32_t)(ent.x >> 32);
    uint4 r4 = make_uint4(0u, 0u, 0u, 0u);  // {xy.lo, xy.hi, mult (0 here), block offset}
    uint2 hdr = make_uint2(0u, 0u), zb = make_uint2(0u, 0u);
    if (active) {
        r4 = lds128(tile_s + k * 16u);
        if (r4.w) {
            hdr = lds64(tile_s + r4.w);
            zb = lds64(tile_s + r4.w + 8u);
        }
    }
    const uint32_t kind = hdr.y & 0xffu, nbits = hdr.y >> 8;
    const uint32_t sign = (uint32_t)__popc((plo & zb.x) ^ (phi & zb.y)) << 31;
    double hr = 0.0, hi = 0.0;
    if (kind == 2u || kind == 4u) {  // A[slot] (+ sum over the occupied positions of x' of D[r][slot] for kind 2)
        const uint32_t slot = (((phi & r4.y) * ENUM_FOLD + (plo & r4.x)) * hdr.x) >> 29;
        const uint32_t a0 = tile_s + r4.w + 16u + slot * 8u, im_delta = (uint32_t)(kind == 2u ? 1 + n : 1) << (nbits + 3u);
        hr = lds_f64(a0);
        if (!REAL) hi = lds_f64(a0 + im_delta);
        if (kind == 2u) {
            for (uint32_t w = plo; w; w &= w - 1u) {
                const uint32_t a = a0 + ((uint32_t)__ffs(w) << (nbits + 3u));  // row 1 + r
                hr += lds_f64(a);
                if (!REAL) hi += lds_f64(a + im_delta);
            }
            for (uint32_t w = phi; w; w &= w - 1u) {
                const uint32_t a = a0 + ((uint32_t)(32 + __ffs(w)) << (nbits + 3u));
                hr += lds_f64(a);
                if (!REAL) hi += lds_f64(a + im_delta);
            }
        }
        hr = flip_hi(hr, sign);
        if (!REAL) hi = flip_hi(hi, sign);
    }
    unsigned dmask = __ballot_sync(0xffffffffu, kind == 3u);
    while (dmask) {  // diagonal: K + sum_i n_i (a_i + sum_{j<i} n_j b_ij), orbitals i over the lanes
        const int src = __ffs(dmask) - 1;
        dmask &= dmask - 1;
        const uint64_t xs = __shfl_sync(0xffffffffu, (unsigned long long)ent.x, src);
        const uint32_t blk = tile_s + __shfl_sync(0xffffffffu, r4.w, src) + 16u;
        const uint32_t im_delta = (uint32_t)(1 + n + n * n) * 8u;
        double pr = 0.0, pi = 0.0;
        for (int i = lane; i < n; i += 32) {
            if (!((xs >> i) & 1ull)) continue;
            double ir = lds_f64(blk + 8u + (uint32_t)i * 8u), ii = REAL ? 0.0 : lds_f64(blk + 8u + (uint32_t)i * 8u + im_delta);
            const uint32_t row = blk + 8u + (uint32_t)n * 8u + (uint32_t)(i * n) * 8u;
            for (uint64_t w = xs & ((1ull << i) - 1ull); w; w &= w - 1ull) {
                const uint32_t a = row + (uint32_t)(__ffsll((long long)w) - 1) * 8u;
                ir += lds_f64(a);
                if (!REAL) ii += lds_f64(a + im_delta);
            }
            pr += ir;
            pi += ii;
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) {
            pr += __shfl_xor_sync(0xffffffffu, pr, dd);
            if (!REAL) pi += __shfl_xor_sync(0xffffffffu, pi, dd);
        }
        if (lane == src) {
            hr = lds_f64(blk) + pr;
            if (!REAL) hi = lds_f64(blk + im_delta) + pi;
        }
    }
    const bool fb = active && kind == 0u;  // neither table applies: term records from global memory
    if (__any_sync(0xffffffffu, fb)) {
        int2 g = make_int2(0, 0);
        if (fb) g = __ldg(t.grp + u0 + k);
        double gr, gi;
        warp_matrix_elements<REAL>(t, fb, g, deinterleave(ent.x), gr, gi);
        if (fb) {
            hr = gr;
            hi = gi;
        }
    }
    if (active) {
        const int64_t r = (int64_t)(ent.y >> 16);
        if (HC == 1) __stcs(H + r, hr);
        if (HC == 2) __stcs(reinterpret_cast<double2 *>(H) + r, make_double2(hr, hi));
    }
}

// One unit of a sample OUTSIDE the (N_alpha, N_beta) sector (never produced by the masked samplers, but legal input): the
// tables do not apply, every matrix element comes from the term arrays in global memory (PO:256-324 as written).
template <bool REAL, int HC>
__device__ __noinline__ void et_unit_slow(const Tables *td, uint32_t tile_s, uint32_t u0, int n_words, uint32_t q_s, int64_t s, uint64_t x,
                                          const uint32_t *row, int64_t out, int32_t *dest, int64_t *xprime, int32_t *xy_ptr, double *H) {
    const Tables &t = *td;
    const int lane = lane_id();
    for (int j = 0; j < n_words; j += ET_STEP_WORDS) {
        const uint32_t w = (j + lane < n_words) ? __ldg(row + j + lane) : 0u;
        int lane_start;
        const int qlen = et_expand(q_s, 0, w, j, lane_start);
        for (int done = 0; done < qlen; done += 32) {
            const bool active = done + lane < qlen;
            const uint32_t k = active ? lds_u16(q_s + 2u * (uint32_t)(done + lane)) : 0u;
            const uint2 m = lds64(tile_s + k * 16u);
            const uint64_t xp = x ^ (((uint64_t)m.y << 32) | m.x);
            double hr = 0.0, hi = 0.0;
            if (HC) {
                int2 g = make_int2(0, 0);
                if (active) g = __ldg(t.grp + u0 + k);
                warp_matrix_elements<REAL>(t, active, g, deinterleave(xp), hr, hi);
            }
            if (active) {
                const int64_t r = out + done + lane;
                if (dest) dest[r] = (int32_t)s;
                xprime[r] = (int64_t)xp;
                if (xy_ptr) xy_ptr[r] = (int32_t)(u0 + k);
                if (HC == 1) H[r] = hr;
                if (HC == 2) reinterpret_cast<double2 *>(H)[r] = make_double2(hr, hi);
            }
        }
        out += qlen;
        __syncwarp();
    }
}

// FLAGS: bit 0 = dest is written, bit 1 = xy_ptr is written
template <bool REAL, int HC, int FLAGS>
__global__ void __launch_bounds__(EN_THREADS, 1)
enum_emit_kernel(Tables t, const int64_t *__restrict__ samples, int64_t n, int alpha, int beta, const uint32_t *__restrict__ bitmap,
                 const int64_t *__restrict__ offsets, const int32_t *__restrict__ tile_prefix, uint32_t *__restrict__ counters,
                 int64_t band_rows, int grab, ulonglong2 *__restrict__ defer_ws, int32_t *__restrict__ dest, int64_t *__restrict__ xprime, int32_t *__restrict__ xy_ptr,
                 double *__restrict__ H) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ EnumTile s_tile;
    __shared__ int s_skip;
    constexpr bool WITH_DEST = (FLAGS & 1) != 0, WITH_PTR = (FLAGS & 2) != 0;
    // lane, warp and the queue address are pinned in registers (volatile asm): left alone, the compiler re-derives them from
    // %tid.x at every use - a tenth of the instructions of this kernel
    int lane, warp;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("shr.u32 %0, %1, 5;" : "=r"(warp) : "r"(threadIdx.x));
    uint32_t q_s = smem_u32(smem_raw) + (uint32_t)warp * ET_QCAP * 2u + 2u * (uint32_t)lane;  // the lane's own slot of the warp's queue
    asm volatile("" : "+r"(q_s));
    unsigned char *tile_buf = smem_raw + ET_QUEUE_BYTES;
    const uint32_t tile_s = smem_u32(tile_buf);
    ulonglong2 *dq = defer_ws + ((size_t)blockIdx.x * EN_WARPS + warp) * ENUM_DEFER_CAP;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;
    const int n_tiles = t.n_enum_tiles;
    // Samples are processed in BANDS of band_rows: all the tiles of one band before the next band.  The twelve or so tile
    // regions of one sample's output rows are then written within a short time of each other (whatever the relative speed of
    // the tiles), so they meet in L2 and go to HBM as long contiguous runs instead of 2-3 KB pieces.
    const int64_t n_bands = (n + band_rows - 1) / band_rows;
    int resident = -1;  // tile in shared memory (the same value in every thread of the CTA)
    for (int64_t band = 0; band < n_bands; ++band)
    for (int tried = 0; tried < n_tiles; ++tried) {
        const int ti = (int)((blockIdx.x + (unsigned)tried) % (unsigned)n_tiles);
        const int64_t band0 = band * band_rows, band_len = min(band_rows, n - band0);
        uint32_t *counter = counters + band * n_tiles + ti;
        __syncthreads();  // every warp is done with the tile that is resident now
        if (threadIdx.x == 0) {
            const uint32_t taken = *reinterpret_cast<volatile uint32_t *>(counter);
            s_skip = (int64_t)taken >= band_len;
            if (!s_skip && ti != resident) {
                const EnumTile et = t.enum_tiles[ti];
                s_tile = et;
                stage_blob(tile_buf, t.enum_blob + et.blob_off, et.blob_bytes, &bar);
            }
        }
        __syncthreads();
        if (s_skip) continue;
        const bool fresh = ti != resident;
        resident = ti;
        const uint32_t u0 = s_tile.u0, im_delta = s_tile.n_tab * 8u, word0 = s_tile.word0, gen_s = tile_s + s_tile.gen_off;
        const int n_words = (int)s_tile.n_words;
        // Units are taken `grab` at a time: lane i of the warp loads the sample, output offset and tile rank of unit i in one
        // round trip, the units are then processed in turn with their values broadcast from the lanes.  The next group's
        // ticket is requested one group ahead and its per-unit values are loaded while the last unit of the current group is
        // processed; the bitmap words of every step are loaded one step ahead, across unit and group boundaries.  Nothing
        // of a unit waits on global memory except the very first group of a tile.
        auto take = [&]() -> uint32_t {  // lane 0 holds the result
            uint32_t v = 0;
            if (lane == 0) v = atomicAdd(counter, (uint32_t)grab);
            return v;
        };
        auto load_meta = [&](uint32_t base, uint64_t &xs, int64_t &os) {
            xs = 0ull;
            os = 0;
            const int64_t sidx = band0 + (int64_t)base + lane;
            if (lane < grab && (int64_t)base + lane < band_len) {
                xs = (uint64_t)__ldg(samples + sidx);
                os = __ldg(offsets + sidx) + __ldg(tile_prefix + sidx * n_tiles + ti);
            }
        };
        uint32_t base = take();
        if (fresh) {
            mbar_wait(&bar, parity);
            parity ^= 1u;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t nbase = take();  // lane 0; read one group later
        uint64_t xs;
        int64_t os;
        load_meta(base, xs, os);
        uint32_t wnext = 0u;
        if ((int64_t)base < band_len && lane < n_words) wnext = __ldg(bitmap + (band0 + (int64_t)base) * t.row_words + word0 + lane);
        int ndef = 0;  // deferred connections of this warp for the resident tile
        while ((int64_t)base < band_len) {
          const int cnt = (int)min((int64_t)grab, band_len - (int64_t)base);
          uint32_t base2 = 0xffffffffu, nbase2 = 0u;
          uint64_t xs2 = 0ull;
          int64_t os2 = 0;
          for (int ui = 0; ui < cnt; ++ui) {
            const int64_t s = band0 + (int64_t)base + ui;
            const bool last_unit = ui == cnt - 1;
            if (last_unit) {  // the next group: its ticket has been in flight for a whole group
                base2 = __shfl_sync(0xffffffffu, nbase, 0);
                if ((int64_t)base2 < band_len) {
                    nbase2 = take();
                    load_meta(base2, xs2, os2);
                }
            }
            const uint64_t x = __shfl_sync(0xffffffffu, (unsigned long long)xs, ui);
            const int64_t out0 = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)os, ui);
            const uint32_t *row = bitmap + s * t.row_words + word0;
            // first words of the unit that follows this one (next unit of the group, or first unit of the next group)
            const uint32_t *row_after = last_unit ? (((int64_t)base2 < band_len) ? bitmap + (band0 + (int64_t)base2) * t.row_words + word0 : nullptr)
                                                  : row + t.row_words;
            if (__popcll(x & 0x5555555555555555ULL) != alpha || __popcll(x & 0xAAAAAAAAAAAAAAAAULL) != beta) {
                et_unit_slow<REAL, HC>(t.dev_copy, tile_s, u0, n_words, q_s - 2u * (uint32_t)lane, s, x, row, out0, WITH_DEST ? dest : nullptr, xprime, WITH_PTR ? xy_ptr : nullptr, H);
                wnext = (row_after && lane < n_words) ? __ldg(row_after + lane) : 0u;
                continue;
            }
            const uint32_t xlo = (uint32_t)x, xhi = (uint32_t)(x >> 32);
            // exclusive prefix parity of the sample: bit i = parity of the bits of x below i (the Jordan-Wigner sign of a
            // pattern connection is parity(S & xy), see analyse_group in abi_core.cu)
            uint64_t S = x;
            S ^= S << 1; S ^= S << 2; S ^= S << 4; S ^= S << 8; S ^= S << 16; S ^= S << 32;
            S <<= 1;
            const uint32_t slo = (uint32_t)S, shi = (uint32_t)(S >> 32);
            // Every warp store covers one 32-row block of the output arrays that starts on a multiple of 32 rows (128 bytes of
            // dest, 256 of x' and H): lane l always writes row (block start + l).  Stores that straddle those blocks, as a
            // unit's natural start out0 would make all of them, run at half the rate (partial 32-byte sectors at both ends of
            // every store, measured in scripts/microbench_write2.cu: 3.5 instead of 6.3 TB/s).  So the first batch of a unit
            // is a partial one on lanes [shift, 32), shift = out0 mod 32, and the last one a partial one on the low lanes.
            const int shift = (int)(out0 & 31);
            EtOut o;
            o.dest = WITH_DEST ? dest + (out0 - shift) + lane : nullptr;
            o.xprime = reinterpret_cast<long long *>(xprime) + (out0 - shift) + lane;
            o.xy_ptr = WITH_PTR ? xy_ptr + (out0 - shift) + lane : nullptr;
            o.H = HC ? H + (int64_t)HC * ((out0 - shift) + lane) : nullptr;
            int emitted = 0;  // connections of this unit already handed to a batch
            int qlen = 0;

            // one batch: lanes [a, a + cnt) take queue entries [from, from + cnt)
            auto batch = [&](int from, int a, int cnt) {
                const bool active = lane >= a && lane < a + cnt;
                const uint32_t k = active ? lds_u16(q_s + 2u * (uint32_t)(from - a)) : 0u;
                const uint4 r4 = lds128(tile_s + k * 16u);
                const uint32_t plo = xlo ^ r4.x, phi = xhi ^ r4.y;
                const uint64_t xp = ((uint64_t)phi << 32) | plo;
                if (active) {
                    if (WITH_DEST) __stcs(o.dest, (int32_t)s);
                    __stcs(o.xprime, (long long)xp);
                    if (WITH_PTR) __stcs(o.xy_ptr, (int32_t)(u0 + k));
                }
                if (WITH_DEST) o.dest += 32;
                o.xprime += 32;
                if (WITH_PTR) o.xy_ptr += 32;
                if (HC) {  // generic connections (mult = 0) were put on the deferred list when their bits were expanded
                    const uint32_t sign = (uint32_t)__popc((slo & r4.x) ^ (shi & r4.y)) << 31;
                    const uint32_t slot = (((phi & r4.y) * ENUM_FOLD + (plo & r4.x)) * r4.z) >> 29;
                    const uint32_t g = tile_s + r4.w + slot * 8u;
                    const double hr = flip_hi(lds_f64(g), sign);
                    if (active && r4.z != 0u) {
                        if (HC == 1) __stcs(o.H, hr);
                        if (HC == 2) __stcs(reinterpret_cast<double2 *>(o.H), make_double2(hr, REAL ? 0.0 : flip_hi(lds_f64(g + im_delta), sign)));
                    }
                    o.H += 32 * HC;
                }
                emitted += cnt;
            };

            for (int j = 0; j < n_words; j += ET_STEP_WORDS) {
                const uint32_t w = wnext;
                if (j + ET_STEP_WORDS < n_words) wnext = (j + ET_STEP_WORDS + lane < n_words) ? __ldg(row + j + ET_STEP_WORDS + lane) : 0u;
                else wnext = (row_after && lane < n_words) ? __ldg(row_after + lane) : 0u;
                int lane_start;
                const int added = et_expand(q_s - 2u * (uint32_t)lane, qlen, w, j, lane_start);
                if (HC) {  // bits of generic masks in this step: onto the deferred list, with their output positions
                    uint32_t gw = (j + lane < n_words) ? (w & lds32(gen_s + 4u * (uint32_t)(j + lane))) : 0u;
                    unsigned dm = __ballot_sync(0xffffffffu, gw != 0u);
                    while (dm) {
                        if (gw) {
                            const uint32_t bit = (uint32_t)__ffs(gw) - 1u;
                            gw &= gw - 1u;
                            const uint32_t k = ((uint32_t)(j + lane) << 5) + bit;
                            const uint2 m = lds64(tile_s + k * 16u);
                            const int64_t p = out0 + emitted + lane_start + __popc(w & ((1u << bit) - 1u));
                            dq[ndef + __popc(dm & lanemask_lt())] =
                                make_ulonglong2((((uint64_t)(xhi ^ m.y)) << 32) | (xlo ^ m.x), ((unsigned long long)p << 16) | k);
                        }
                        ndef += __popc(dm);
                        __syncwarp();
                        if (ndef >= 32) {
                            const ulonglong2 ent = __ldcg(dq + lane);
                            et_eval_deferred<REAL, HC>(t.dev_copy, tile_s, u0, ent, true, H);
                            const ulonglong2 mv = (32 + lane < ndef) ? __ldcg(dq + 32 + lane) : make_ulonglong2(0ull, 0ull);
                            __syncwarp();
                            if (32 + lane < ndef) dq[lane] = mv;
                            ndef -= 32;
                            __syncwarp();
                        }
                        dm = __ballot_sync(0xffffffffu, gw != 0u);
                    }
                }
                qlen += added;
                int done = 0, a = (emitted + shift) & 31;  // a != 0 only before the first batch of the unit
                while (qlen - done >= 32 - a) {
                    batch(done, a, 32 - a);
                    done += 32 - a;
                    a = 0;
                }
                if (done > 0) {  // move the remainder (< 32 entries) to the front of the queue
                    const int rem = qlen - done;
                    const uint32_t v = lane < rem ? lds_u16(q_s + 2u * (uint32_t)done) : 0u;
                    __syncwarp();
                    if (lane < rem) sts_u16(q_s, v);
                    __syncwarp();
                    qlen = rem;
                }
            }
            if (qlen > 0) batch(0, (emitted + shift) & 31, qlen);
            __syncwarp();
          }
          base = base2;
          nbase = nbase2;
          xs = xs2;
          os = os2;
        }
        if (HC && ndef > 0) {  // what is left of the deferred list before the tile goes away
            const ulonglong2 ent = lane < ndef ? __ldcg(dq + lane) : make_ulonglong2(0ull, 0ull);
            et_eval_deferred<REAL, HC>(t.dev_copy, tile_s, u0, ent, lane < ndef, H);
        }
    }
}

static int filter_warps(const Tables *t) {
    const int64_t avail = (int64_t)EN_SMEM_MAX - 1024 - t->tile_bytes_max;
    int64_t w = avail / (t->row_words * 4);
    if (w > EN_WARPS) w = EN_WARPS;
    return (int)(w & ~3ll);
}

// bitmap words of the 32 samples' rows kept in shared memory at a time: the whole row when it fits beside a second CTA
constexpr int BS_CHUNK_MAX = 894;  // 32 * 895 * 4 B = 112 KB
static int bitsliced_chunk(const Tables *t) {
    const int64_t nchunks = (t->row_words + BS_CHUNK_MAX - 1) / BS_CHUNK_MAX;
    return (int)((t->row_words + nchunks - 1) / nchunks);
}
static size_t bitsliced_smem(const Tables *t) { return (size_t)32 * (size_t)(bitsliced_chunk(t) | 1) * 4; }
static bool bitsliced_available(const Tables *t) { return t->bs_ok != 0; }

static bool tiled_available(const Tables *t) {
    return t->n_enum_tiles > 0 && (bitsliced_available(t) || filter_warps(t) >= 4) &&
           ET_QUEUE_BYTES + t->enum_tile_bytes_max + 256 <= EN_SMEM_MAX;
}

// rows per band of the emit kernel (see enum_emit_kernel); ANQS_ENUM_BAND_ROWS overrides it for experiments
static int64_t emit_band_rows() {
    static int64_t cached = 0;
    if (cached == 0) {
        const char *e = getenv("ANQS_ENUM_BAND_ROWS");
        const long long v = e ? atoll(e) : 0;
        cached = v >= 32 ? (int64_t)v : ((int64_t)1 << 40);
    }
    return cached;
}
static int64_t emit_bands(int64_t n) { return std::max<int64_t>(1, (n + emit_band_rows() - 1) / emit_band_rows()); }
static size_t counters_bytes(const Tables *t, int64_t n) {
    return ((size_t)emit_bands(n) * (size_t)std::max(1, t->n_enum_tiles) * 4 + 127) / 128 * 128;
}
static size_t prefix_bytes(const Tables *t, int64_t n) {
    return ((size_t)n * (size_t)std::max(1, t->n_enum_tiles) * sizeof(int32_t) + 127) / 128 * 128;
}
static size_t defer_bytes() { return (size_t)sm_count_of_current_device() * EN_WARPS * ENUM_DEFER_CAP * sizeof(ulonglong2); }

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_k1_enum_tiles(const anqs_tables_t *h) {
    if (!h) return 0;
    const Tables *t = (const Tables *)h;
    return tiled_available(t) ? t->n_enum_tiles : 0;
}

size_t anqs_k1_enum_workspace(const anqs_tables_t *h, int64_t n) {
    if (!h || n < 0) return 0;
    const Tables *t = (const Tables *)h;
    // [tile counters][tile_prefix: n x n_tiles int32][deferred lists of the emit kernel: one per warp]
    return counters_bytes(t, n) + prefix_bytes(t, n) + defer_bytes();
}

int anqs_k1_enum_filter(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                        int64_t *d_counts, uint32_t *d_bitmap, void *d_work, void *stream) {
    return anqs_k1_enum_filter_variant(h, d_samples, n, alpha_num, beta_num, d_counts, d_bitmap, d_work, 0, stream);
}

int anqs_k1_enum_filter_variant(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                                int64_t *d_counts, uint32_t *d_bitmap, void *d_work, int variant, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (automatic) or 1 (product-layout filter)");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    const Tables *t = (const Tables *)h;
    ANQS_REQUIRE(tiled_available(t), "the tiled enumeration is unavailable for this table (anqs_k1_enum_tiles() == 0): use anqs_k1_filter / anqs_k1_emit");
    ANQS_REQUIRE(d_samples && d_counts && d_bitmap && d_work, "null pointer");
    if (bitsliced_available(t) && variant != 1) {
        const size_t smem = bitsliced_smem(t);
        ANQS_CUDA(cudaFuncSetAttribute(enum_filter_bitsliced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(EN_SMEM_MAX - 1024) / (smem + 1024)));
        const int grid = (int)std::min<int64_t>((n + 31) / 32, (int64_t)sm_count_of_current_device() * per_sm);
        int32_t *tile_prefix = reinterpret_cast<int32_t *>((unsigned char *)d_work + counters_bytes(t, n));
        enum_filter_bitsliced_kernel<<<grid, BS_THREADS, smem, (cudaStream_t)stream>>>(*t, d_samples, n, alpha_num, beta_num, d_counts,
                                                                                       d_bitmap, tile_prefix, bitsliced_chunk(t));
        ANQS_LAUNCH_CHECK();
        return 0;
    }
    const int warps = filter_warps(t);
    const uint32_t bm_bytes = (uint32_t)(((size_t)warps * t->row_words * 4 + 127) / 128 * 128);
    const size_t smem = (size_t)bm_bytes + (size_t)t->tile_bytes_max;
    ANQS_CUDA(cudaFuncSetAttribute(enum_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ngroups = (n + warps - 1) / warps;
    const int grid = (int)std::min<int64_t>(ngroups, sm_count_of_current_device());
    int32_t *tile_prefix = reinterpret_cast<int32_t *>((unsigned char *)d_work + counters_bytes(t, n));
    enum_filter_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(*t, d_samples, n, alpha_num, beta_num, d_counts, d_bitmap,
                                                                          tile_prefix, bm_bytes);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_k1_enum_emit(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                      const uint32_t *d_bitmap, const int64_t *d_offsets, void *d_work, int32_t *d_dest, int64_t *d_xprime, int32_t *d_xy_ptr,
                      double *d_H, int h_components, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(n < ((int64_t)1 << 31), "chunk too large for int32 dest; split the batch");
    const Tables *t = (const Tables *)h;
    ANQS_REQUIRE(tiled_available(t), "the tiled enumeration is unavailable for this table (anqs_k1_enum_tiles() == 0): use anqs_k1_filter / anqs_k1_emit");
    ANQS_REQUIRE(d_samples && d_bitmap && d_offsets && d_work && d_xprime, "null pointer");
    const int hc = d_H ? h_components : 0;
    ANQS_REQUIRE(hc == 0 || hc == 1 || hc == 2, "h_components must be 1 (real) or 2 (complex)");
    ANQS_REQUIRE(!(hc == 1 && !t->weights_real), "real matrix elements requested but the Hamiltonian weights are complex");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *counters = reinterpret_cast<uint32_t *>(d_work);
    const int32_t *tile_prefix = reinterpret_cast<const int32_t *>((unsigned char *)d_work + counters_bytes(t, n));
    ANQS_CUDA(cudaMemsetAsync(counters, 0, counters_bytes(t, n), s));
    ulonglong2 *defer_ws = reinterpret_cast<ulonglong2 *>((unsigned char *)d_work + counters_bytes(t, n) + prefix_bytes(t, n));
    const size_t smem = (size_t)ET_QUEUE_BYTES + (size_t)t->enum_tile_bytes_max;
    const int grid = sm_count_of_current_device();
    const int flags = (d_dest ? 1 : 0) | (d_xy_ptr ? 2 : 0);
    // units a warp takes per ticket (<= 32): about a sixth of its share of one band of one tile, so the tail stays short
    int grab = 1;
    {
        const int64_t warps_per_tile = std::max<int64_t>(1, (int64_t)grid * EN_WARPS / std::max(1, t->n_enum_tiles));
        const int64_t share = std::min<int64_t>(n, emit_band_rows()) / (warps_per_tile * 6);
        while (grab * 2 <= share && grab < 32) grab *= 2;
        if (const char *e = getenv("ANQS_ENUM_GRAB")) grab = std::max(1, std::min(32, atoi(e)));
    }
#define ANQS_ENUM_EMIT_F(REAL, HC, FLAGS)                                                                                     \
    do {                                                                                                                      \
        auto kern = enum_emit_kernel<REAL, HC, FLAGS>;                                                                        \
        ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
        kern<<<grid, EN_THREADS, smem, s>>>(*t, d_samples, n, alpha_num, beta_num, d_bitmap, d_offsets, tile_prefix, counters, emit_band_rows(), grab, defer_ws, d_dest,   \
                                            d_xprime, d_xy_ptr, d_H);                                                         \
    } while (0)
#define ANQS_ENUM_EMIT(REAL, HC)                                                                                              \
    do {                                                                                                                      \
        if (flags == 0) ANQS_ENUM_EMIT_F(REAL, HC, 0);                                                                        \
        else if (flags == 1) ANQS_ENUM_EMIT_F(REAL, HC, 1);                                                                   \
        else if (flags == 2) ANQS_ENUM_EMIT_F(REAL, HC, 2);                                                                   \
        else ANQS_ENUM_EMIT_F(REAL, HC, 3);                                                                                   \
    } while (0)
    if (hc == 0) ANQS_ENUM_EMIT(true, 0);
    else if (t->weights_real && hc == 1) ANQS_ENUM_EMIT(true, 1);
    else if (t->weights_real && hc == 2) ANQS_ENUM_EMIT(true, 2);
    else ANQS_ENUM_EMIT(false, 2);
#undef ANQS_ENUM_EMIT
#undef ANQS_ENUM_EMIT_F
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
