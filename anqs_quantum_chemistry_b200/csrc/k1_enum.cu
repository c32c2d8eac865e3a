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
                // connections of every enumeration tile that overlaps this chunk (raw counts for now, scanned below)
                for (int e = e_first; e < n_tiles; ++e) {
                    const int w0 = (int)__ldg(&t.enum_tiles[e].word0), w1 = w0 + (int)__ldg(&t.enum_tiles[e].n_words);
                    if (w0 >= c1) break;
                    int c = 0;
                    for (int j = max(w0, c0) + lane; j < min(w1, c1); j += 32) c += __popc(bm[j]);
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == 0) {
                        int32_t *slot = tile_prefix + r * n_tiles + e;
                        *slot = w0 >= c0 ? c : *slot + c;  // a tile that started in an earlier chunk already has a partial count
                    }
                }
                uint32_t *row = bitmap + r * t.row_words;
                for (int j = c0 + lane; j < c1; j += 32) row[j] = bm[j];
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
constexpr int ET_STEP_WORDS = ENUM_STEP_WORDS;
constexpr int ET_QCAP = ENUM_QCAP;
constexpr int ET_QUEUE_BYTES = ENUM_QUEUE_BYTES;
static_assert(EN_WARPS == ENUM_EMIT_WARPS, "queue budget");
constexpr uint32_t ET_BIG = 24;                   // YZ groups longer than this are summed by the whole warp

struct EtTile {
    const uint64_t *xy, *zb;
    const uint2 *desc;
    const double *tab;   // pattern tables: [n_tab] re, then [n_tab] im when weights are complex
    const uint4 *term;   // {yz.lo, yz.hi, w.lo, w.hi}
    const double *wim;
    uint32_t u0, n_tab;
};

__device__ __forceinline__ double flip_hi(double w, uint32_t sign) {
    return __hiloint2double(__double2hiint(w) ^ (int)sign, __double2loint(w));
}

template <bool REAL>
__device__ __forceinline__ void et_term(const uint4 rec, const double *wim, uint32_t k, uint32_t xlo, uint32_t xhi, double &hr,
                                        double &hi) {
    const uint32_t sign = (uint32_t)__popc((xlo & rec.x) ^ (xhi & rec.y)) << 31;
    hr += __hiloint2double((int)(rec.w ^ sign), (int)rec.z);
    if (!REAL) hi += flip_hi(wim[k], sign);
}

// One connection per lane: x' = x ^ xy[u], H_{x,x'} and the stores.  Must be called by the whole warp.
template <bool REAL, int HC>
__device__ __forceinline__ void et_emit_batch(const Tables &t, const EtTile &tl, uint64_t x, bool in_sector, int s, bool active,
                                              uint32_t ul, int64_t r, int32_t *dest, int64_t *xprime, int32_t *xy_ptr, double *H) {
    uint64_t xp = 0;
    double hr = 0.0, hi = 0.0;
    if (active) xp = x ^ tl.xy[ul];
    if (HC && in_sector) {
        uint2 d = make_uint2(0u, 0u);
        uint64_t zb = 0;
        if (active) {
            d = tl.desc[ul];
            zb = tl.zb[ul];
        }
        const uint32_t xlo = (uint32_t)xp, xhi = (uint32_t)(xp >> 32);
        const uint32_t nbits = d.y & 3u;
        const uint32_t num = nbits ? 0u : d.y >> 2;
        if (nbits) {
            // pattern group: sign from the Z string outside the mask, magnitude from the table
            const uint32_t sign = (uint32_t)__popc((xlo & (uint32_t)zb) ^ (xhi & (uint32_t)(zb >> 32))) << 31;
            uint32_t idx = (uint32_t)(xp >> ((d.y >> 2) & 63u)) & 1u;
            idx |= ((uint32_t)(xp >> ((d.y >> 8) & 63u)) & 1u) << 1;
            idx |= ((uint32_t)(xp >> ((d.y >> 14) & 63u)) & 1u) << 2;
            idx = (idx & ((1u << nbits) - 1u)) + d.x;
            hr = flip_hi(tl.tab[idx], sign);
            if (!REAL) hi = flip_hi(tl.tab[tl.n_tab + idx], sign);
        } else if (num <= ET_BIG) {
            double h2 = 0.0, i2 = 0.0;
            uint32_t k = d.x;
            const uint32_t end = d.x + num;
            for (; k + 1 < end; k += 2) {
                et_term<REAL>(tl.term[k], tl.wim, k, xlo, xhi, hr, hi);
                et_term<REAL>(tl.term[k + 1], tl.wim, k + 1, xlo, xhi, h2, i2);
            }
            if (k < end) et_term<REAL>(tl.term[k], tl.wim, k, xlo, xhi, hr, hi);
            hr += h2;
            hi += i2;
        }
        unsigned bigmask = __ballot_sync(0xffffffffu, num > ET_BIG);
        while (bigmask) {  // long generic groups (the diagonal, one-body excitations): the whole warp sums one group
            const int src = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const uint32_t st = __shfl_sync(0xffffffffu, d.x, src), end = st + __shfl_sync(0xffffffffu, num, src);
            const uint32_t slo = __shfl_sync(0xffffffffu, xlo, src), shi = __shfl_sync(0xffffffffu, xhi, src);
            double sr = 0.0, si = 0.0;
            for (uint32_t k = st + lane_id(); k < end; k += 32) et_term<REAL>(tl.term[k], tl.wim, k, slo, shi, sr, si);
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, dd);
                if (!REAL) si += __shfl_xor_sync(0xffffffffu, si, dd);
            }
            if (lane_id() == src) {
                hr = sr;
                hi = si;
            }
        }
    } else if (HC) {
        // sample outside the (N_alpha, N_beta) sector: the pattern tables do not apply; PO:256-324 on the untiled term arrays
        int2 g = make_int2(0, 0);
        if (active) g = __ldg(t.grp + tl.u0 + ul);
        warp_matrix_elements<REAL>(t, active, g, deinterleave(xp), hr, hi);
    }
    if (active) {
        if (dest) __stcs(dest + r, s);
        __stcs(reinterpret_cast<long long *>(xprime) + r, (long long)xp);
        if (xy_ptr) __stcs(xy_ptr + r, (int32_t)(tl.u0 + ul));
        if (HC == 1) __stcs(H + r, hr);
        if (HC == 2) __stcs(reinterpret_cast<double2 *>(H) + r, make_double2(hr, hi));
    }
}

template <bool REAL, int HC>
__global__ void __launch_bounds__(EN_THREADS, 1)
enum_emit_kernel(Tables t, const int64_t *__restrict__ samples, int64_t n, int alpha, int beta, const uint32_t *__restrict__ bitmap,
                 const int64_t *__restrict__ offsets, const int32_t *__restrict__ tile_prefix, uint32_t *__restrict__ counters,
                 int32_t *__restrict__ dest, int64_t *__restrict__ xprime, int32_t *__restrict__ xy_ptr, double *__restrict__ H) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ EnumTile s_tile;
    __shared__ int s_skip;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint16_t *q = reinterpret_cast<uint16_t *>(smem_raw) + warp * ET_QCAP;
    unsigned char *tile_buf = smem_raw + ET_QUEUE_BYTES;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    uint32_t parity = 0;
    const int n_tiles = t.n_enum_tiles;
    int ti = (int)(blockIdx.x % (unsigned)n_tiles);
    for (int tried = 0; tried < n_tiles; ++tried, ti = (ti + 1 == n_tiles) ? 0 : ti + 1) {
        __syncthreads();  // every warp is done with the tile that is resident now
        if (threadIdx.x == 0) {
            const uint32_t taken = *reinterpret_cast<volatile uint32_t *>(counters + ti);
            s_skip = taken >= (uint64_t)n;
            if (!s_skip) {
                const EnumTile et = t.enum_tiles[ti];
                s_tile = et;
                stage_blob(tile_buf, t.enum_blob + et.blob_off, et.blob_bytes, &bar);
            }
        }
        __syncthreads();
        if (s_skip) continue;
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const EnumTile et = s_tile;
        EtTile tl;
        tl.xy = reinterpret_cast<const uint64_t *>(tile_buf);
        tl.zb = tl.xy + et.n_masks;
        tl.desc = reinterpret_cast<const uint2 *>(tl.zb + et.n_masks);
        tl.tab = reinterpret_cast<const double *>(tile_buf + et.tab_off);
        tl.term = reinterpret_cast<const uint4 *>(tile_buf + et.term_off);
        tl.wim = reinterpret_cast<const double *>(tile_buf + et.term_off + (size_t)et.n_terms * 16);
        tl.u0 = et.u0;
        tl.n_tab = et.n_tab;
        const int n_words = (int)et.n_words;
        for (;;) {
            uint32_t s32 = 0;
            if (lane == 0) s32 = atomicAdd(counters + ti, 1u);
            s32 = __shfl_sync(0xffffffffu, s32, 0);
            if ((int64_t)s32 >= n) break;
            const int64_t s = (int64_t)s32;
            const uint64_t x = (uint64_t)samples[s];
            const bool in_sector = __popcll(x & 0x5555555555555555ULL) == alpha && __popcll(x & 0xAAAAAAAAAAAAAAAAULL) == beta;
            int64_t out = offsets[s] + tile_prefix[s * n_tiles + ti];
            const uint32_t *row = bitmap + s * t.row_words + et.word0;
            int qlen = 0;
            uint32_t wnext = lane < n_words ? __ldg(row + lane) : 0u;
            for (int j = 0; j < n_words; j += 32) {
                const uint32_t wfull = wnext;
                wnext = (j + 32 + lane < n_words) ? __ldg(row + j + 32 + lane) : 0u;
                {
                    uint32_t w = wfull;
                    const int c = __popc(w);
                    int inc = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, inc, d);
                        if (lane >= d) inc += o;
                    }
                    const int total = __shfl_sync(0xffffffffu, inc, 31);
                    if (total == 0) continue;
                    uint16_t *qp = q + (qlen + inc - c);
                    uint32_t base = (uint32_t)(j + lane) << 5;
                    while (w) {
                        const int bit = __ffs(w) - 1;
                        w &= w - 1;
                        *qp++ = (uint16_t)(base + bit);
                    }
                    __syncwarp();
                    qlen += total;
                    int done = 0;
                    while (qlen - done >= 32) {
                        et_emit_batch<REAL, HC>(t, tl, x, in_sector, (int)s, true, q[done + lane], out + done + lane, dest, xprime, xy_ptr, H);
                        done += 32;
                    }
                    if (done > 0) {
                        const int rem = qlen - done;
                        const uint16_t v = lane < rem ? q[done + lane] : (uint16_t)0;
                        __syncwarp();
                        if (lane < rem) q[lane] = v;
                        __syncwarp();
                        out += done;
                        qlen = rem;
                    }
                }
            }
            if (qlen > 0)
                et_emit_batch<REAL, HC>(t, tl, x, in_sector, (int)s, lane < qlen, lane < qlen ? q[lane] : (uint16_t)0, out + lane, dest, xprime,
                                        xy_ptr, H);
            __syncwarp();
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

static bool g_force_product_filter = false;  // test hook (anqs_k1_enum_force_product_filter)

static size_t counters_bytes(const Tables *t) { return ((size_t)t->n_enum_tiles * 4 + 127) / 128 * 128; }

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_k1_enum_tiles(const anqs_tables_t *h) {
    if (!h) return 0;
    const Tables *t = (const Tables *)h;
    return tiled_available(t) ? t->n_enum_tiles : 0;
}

void anqs_k1_enum_force_product_filter(int on) { g_force_product_filter = on != 0; }

size_t anqs_k1_enum_workspace(const anqs_tables_t *h, int64_t n) {
    if (!h || n < 0) return 0;
    const Tables *t = (const Tables *)h;
    return counters_bytes(t) + (size_t)n * (size_t)std::max(1, t->n_enum_tiles) * sizeof(int32_t);
}

int anqs_k1_enum_filter(const anqs_tables_t *h, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                        int64_t *d_counts, uint32_t *d_bitmap, void *d_work, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    const Tables *t = (const Tables *)h;
    ANQS_REQUIRE(tiled_available(t), "the tiled enumeration is unavailable for this table (anqs_k1_enum_tiles() == 0): use anqs_k1_filter / anqs_k1_emit");
    ANQS_REQUIRE(d_samples && d_counts && d_bitmap && d_work, "null pointer");
    if (bitsliced_available(t) && !g_force_product_filter) {
        const size_t smem = bitsliced_smem(t);
        ANQS_CUDA(cudaFuncSetAttribute(enum_filter_bitsliced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(EN_SMEM_MAX - 1024) / (smem + 1024)));
        const int grid = (int)std::min<int64_t>((n + 31) / 32, (int64_t)sm_count_of_current_device() * per_sm);
        int32_t *tile_prefix = reinterpret_cast<int32_t *>((unsigned char *)d_work + counters_bytes(t));
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
    int32_t *tile_prefix = reinterpret_cast<int32_t *>((unsigned char *)d_work + counters_bytes(t));
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
    const int32_t *tile_prefix = reinterpret_cast<const int32_t *>((unsigned char *)d_work + counters_bytes(t));
    ANQS_CUDA(cudaMemsetAsync(counters, 0, counters_bytes(t), s));
    const size_t smem = (size_t)ET_QUEUE_BYTES + (size_t)t->enum_tile_bytes_max;
    const int grid = sm_count_of_current_device();
#define ANQS_ENUM_EMIT(REAL, HC)                                                                                              \
    do {                                                                                                                      \
        auto kern = enum_emit_kernel<REAL, HC>;                                                                               \
        ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
        kern<<<grid, EN_THREADS, smem, s>>>(*t, d_samples, n, alpha_num, beta_num, d_bitmap, d_offsets, tile_prefix, counters, d_dest, d_xprime,   \
                                            d_xy_ptr, d_H);                                                                   \
    } while (0)
    if (hc == 0) ANQS_ENUM_EMIT(true, 0);
    else if (t->weights_real && hc == 1) ANQS_ENUM_EMIT(true, 1);
    else if (t->weights_real && hc == 2) ANQS_ENUM_EMIT(true, 2);
    else ANQS_ENUM_EMIT(false, 2);
#undef ANQS_ENUM_EMIT
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
