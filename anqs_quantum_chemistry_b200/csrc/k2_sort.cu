// Kernel family 2, ordering part: sort / unique / top-k of packed configurations and of Gumbel keys, hand-written
// (reference: HilbertSpace.sort_base_idx HS:239-261, compute_unique_indices HS:215-228, the per-level
// `sort(descending=True)[:sample_num]` of the Gumbel sampler ANQS:733, the de-duplication of non-sampled connected
// configurations PO:1016-1040).  The reference calls torch.sort / torch.unique; here:
//
//   anqs_sort_pairs_u64   stable LSD radix sort of (64-bit key, 64-bit payload) pairs on 8-bit digits, only over the bit
//                         range the caller says can differ (a 20-qubit configuration takes 3 passes, not 8).  Per pass:
//                         per-block digit counts -> one scan -> stable scatter (ranks inside a warp from __match_any_sync).
//   anqs_unique_i64       sort (signed order, payload = original position) -> head flags -> scan -> unique values + inverse.
//   anqs_topk_f64         the k largest of n doubles in descending order, ties by position (= the first k rows of a stable
//                         descending sort): 8-bit radix SELECT from the top digit down (histograms only, no data movement,
//                         no host read), ordered compaction of the survivors, then the sort above on k pairs instead of n.
//
// All HBM-bound: 16 B read + 16 B written per pair and pass for the sort, 8 B read per key and digit for the selection.
#include <algorithm>

#include "common.cuh"

namespace anqs {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32;
constexpr int RS_PER_WARP = 512;                       // elements a warp scatters, in 16 chunks of 32, in order
constexpr int RS_TILE = RS_WARPS * RS_PER_WARP;        // elements per block
constexpr int RS_SCAN_THREADS = 1024;

// key transforms: 0 = unsigned integer as is, 1 = IEEE double bits -> unsigned with the same order
__device__ __forceinline__ uint64_t sort_key(uint64_t bits, int kind, uint64_t xor_mask) {
    if (kind == 1) bits ^= (bits >> 63) ? ~0ull : (1ull << 63);
    return bits ^ xor_mask;
}
__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift, uint32_t mask) { return (uint32_t)(key >> shift) & mask; }

__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const uint64_t *__restrict__ keys, int64_t n, int kind, uint64_t xor_mask, int shift,
                                                              uint32_t mask, uint32_t *__restrict__ counts, int nblocks) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const int64_t j = base + i;
        if (j < n) atomicAdd(&hist[digit_of(sort_key(keys[j], kind, xor_mask), shift, mask)], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];  // digit-major: one scan gives the global offsets
}

// exclusive scan of `m` uint32 counts in place (single block; m <= 256 * blocks, a few 10^5 at most)
__global__ void __launch_bounds__(RS_SCAN_THREADS) rs_scan_kernel(uint32_t *__restrict__ counts, int64_t m) {
    __shared__ uint32_t part[RS_SCAN_THREADS];
    const int64_t per = (m + RS_SCAN_THREADS - 1) / RS_SCAN_THREADS;
    const int64_t lo = min((int64_t)threadIdx.x * per, m), hi = min(lo + per, m);
    uint32_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += counts[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < RS_SCAN_THREADS; d <<= 1) {
        const uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (int64_t i = lo; i < hi; ++i) {
        const uint32_t c = counts[i];
        counts[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t *__restrict__ keys_in, const int64_t *__restrict__ vals_in,
                                                                uint64_t *__restrict__ keys_out, int64_t *__restrict__ vals_out, int64_t n, int kind,
                                                                uint64_t xor_mask, int shift, uint32_t mask, const uint32_t *__restrict__ offsets,
                                                                int nblocks) {
    __shared__ uint32_t wh[RS_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0u;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_PER_WARP;
    // (a) digit counts of this warp's run of elements
    for (int c = 0; c < RS_PER_WARP; c += 32) {
        const int64_t j = wbase + c + lane;
        const uint32_t d = j < n ? digit_of(sort_key(keys_in[j], kind, xor_mask), shift, mask) : 256u + (uint32_t)lane;
        const unsigned m = __match_any_sync(0xffffffffu, d);
        if (j < n && (m & ((1u << lane) - 1u)) == 0u) wh[warp][d] += (uint32_t)__popc(m);  // lowest lane of each digit class
        __syncwarp();
    }
    __syncthreads();
    // (b) per digit: global offset of the block, then the warps of the block in order
    {
        const int d = threadIdx.x;
        uint32_t run = offsets[(size_t)d * nblocks + blockIdx.x];
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // (c) stable scatter, 32 elements at a time
    for (int c = 0; c < RS_PER_WARP; c += 32) {
        const int64_t j = wbase + c + lane;
        const bool ok = j < n;
        const uint64_t bits = ok ? keys_in[j] : 0ull;
        const uint32_t d = ok ? digit_of(sort_key(bits, kind, xor_mask), shift, mask) : 256u + (uint32_t)lane;
        const unsigned m = __match_any_sync(0xffffffffu, d);
        const uint32_t below = (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (ok) {
            const uint32_t pos = wh[warp][d] + below;
            keys_out[pos] = bits;
            vals_out[pos] = vals_in ? vals_in[j] : j;
        }
        __syncwarp();
        if (ok && below == 0u) wh[warp][d] += (uint32_t)__popc(m);
        __syncwarp();
    }
}

// Small inputs (n <= RS_SMALL_MAX): every pass in ONE launch of ONE CTA - counts, scan and stable scatter separated by
// __syncthreads(), ping-pong between the caller's output and the workspace.  A VMC iteration at 1e4 samples sorts a handful of
// such arrays; at three launches per pass they would cost more in launch latency than all its other kernels together.
constexpr int RS_SMALL_THREADS = 1024, RS_SMALL_WARPS = RS_SMALL_THREADS / 32, RS_SMALL_MAX = 1 << 14;

__global__ void __launch_bounds__(RS_SMALL_THREADS) rs_small_kernel(const uint64_t *__restrict__ keys_in, const int64_t *__restrict__ vals_in,
                                                                    uint64_t *keys_out, int64_t *vals_out, uint64_t *tmp_k, int64_t *tmp_v, int n,
                                                                    int begin_bit, int end_bit, int kind, uint64_t xor_mask) {
    __shared__ uint32_t wh[RS_SMALL_WARPS][256];   // per-warp digit bases of the chunk being scattered
    __shared__ uint32_t base[256];                 // running global base per digit
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npass = (end_bit - begin_bit + 7) / 8;
    const uint64_t *src_k = keys_in;
    const int64_t *src_v = vals_in;
    for (int p = 0; p < npass; ++p) {
        const bool to_out = ((npass - 1 - p) & 1) == 0;
        uint64_t *dst_k = to_out ? keys_out : tmp_k;
        int64_t *dst_v = to_out ? vals_out : tmp_v;
        const int shift = begin_bit + 8 * p;
        const uint32_t mask = (uint32_t)((1u << min(8, end_bit - shift)) - 1u);
        // digit histogram of the whole array -> exclusive scan = first output position of every digit
        if (threadIdx.x < 256) base[threadIdx.x] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += RS_SMALL_THREADS) atomicAdd(&base[digit_of(sort_key(src_k[i], kind, xor_mask), shift, mask)], 1u);
        __syncthreads();
        if (warp == 0) {  // 256 bins, 8 per lane
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = base[lane * 8 + i]; sum += v[i]; }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int i = 0; i < 8; ++i) { base[lane * 8 + i] = run; run += v[i]; }
        }
        __syncthreads();
        // stable scatter, 1024 elements (32 per warp) at a time, in order
        for (int c0 = 0; c0 < n; c0 += RS_SMALL_THREADS) {
            for (int i = threadIdx.x; i < RS_SMALL_WARPS * 256; i += RS_SMALL_THREADS) (&wh[0][0])[i] = 0u;
            __syncthreads();
            const int j = c0 + threadIdx.x;
            const bool ok = j < n;
            const uint64_t bits = ok ? src_k[j] : 0ull;
            const uint32_t d = ok ? digit_of(sort_key(bits, kind, xor_mask), shift, mask) : 256u + (uint32_t)lane;
            const unsigned m = __match_any_sync(0xffffffffu, d);
            const uint32_t below = (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (ok && below == 0u) wh[warp][d] = (uint32_t)__popc(m);
            __syncthreads();
            if (threadIdx.x < 256) {  // per digit: this chunk's warps in order behind the running base
                uint32_t run = base[threadIdx.x];
                for (int w = 0; w < RS_SMALL_WARPS; ++w) {
                    const uint32_t c = wh[w][threadIdx.x];
                    wh[w][threadIdx.x] = run;
                    run += c;
                }
                base[threadIdx.x] = run;
            }
            __syncthreads();
            if (ok) {
                const uint32_t pos = wh[warp][d] + below;
                dst_k[pos] = bits;
                dst_v[pos] = src_v ? src_v[j] : (int64_t)j;
            }
            __syncthreads();
        }
        src_k = dst_k;
        src_v = dst_v;
        __syncthreads();
    }
}

__global__ void copy_pairs_kernel(const uint64_t *__restrict__ keys_in, const int64_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                                  int64_t *__restrict__ vals_out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys_out[i] = keys_in[i];
        vals_out[i] = vals_in ? vals_in[i] : i;
    }
}

static int64_t rs_blocks(int64_t n) { return std::max<int64_t>(1, (n + RS_TILE - 1) / RS_TILE); }
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }
// workspace: [ping-pong keys n x 8][ping-pong payloads n x 8][counts 256 x blocks x 4]
static size_t sort_ws_bytes(int64_t n) { return 2 * align256((size_t)std::max<int64_t>(n, 1) * 8) + align256((size_t)256 * rs_blocks(n) * 4); }

static int sort_pairs(const uint64_t *keys_in, const int64_t *vals_in, uint64_t *keys_out, int64_t *vals_out, int64_t n, int begin_bit,
                      int end_bit, int kind, uint64_t xor_mask, void *work, cudaStream_t s) {
    if (n == 0) return 0;
    const int npass = (end_bit - begin_bit + 7) / 8;
    const int grid_copy = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    if (npass == 0) {
        copy_pairs_kernel<<<grid_copy, 256, 0, s>>>(keys_in, vals_in, keys_out, vals_out, n);
        return cudaGetLastError() == cudaSuccess ? 0 : 2;
    }
    unsigned char *w = (unsigned char *)work;
    uint64_t *tmp_k = (uint64_t *)w;
    int64_t *tmp_v = (int64_t *)(w + align256((size_t)n * 8));
    uint32_t *counts = (uint32_t *)(w + 2 * align256((size_t)n * 8));
    if (n <= RS_SMALL_MAX) {
        rs_small_kernel<<<1, RS_SMALL_THREADS, 0, s>>>(keys_in, vals_in, keys_out, vals_out, tmp_k, tmp_v, (int)n, begin_bit, end_bit, kind, xor_mask);
        return cudaGetLastError() == cudaSuccess ? 0 : 2;
    }
    const int nblocks = (int)rs_blocks(n);
    const uint64_t *src_k = keys_in;
    const int64_t *src_v = vals_in;
    for (int p = 0; p < npass; ++p) {
        // the last pass must land in the caller's buffers: passes alternate out / tmp backwards from there
        const bool to_out = ((npass - 1 - p) & 1) == 0;
        uint64_t *dst_k = to_out ? keys_out : tmp_k;
        int64_t *dst_v = to_out ? vals_out : tmp_v;
        const int shift = begin_bit + 8 * p;
        const uint32_t mask = (uint32_t)((1u << std::min(8, end_bit - shift)) - 1u);
        rs_count_kernel<<<nblocks, RS_THREADS, 0, s>>>(src_k, n, kind, xor_mask, shift, mask, counts, nblocks);
        rs_scan_kernel<<<1, RS_SCAN_THREADS, 0, s>>>(counts, (int64_t)256 * nblocks);
        rs_scatter_kernel<<<nblocks, RS_THREADS, 0, s>>>(src_k, src_v, dst_k, dst_v, n, kind, xor_mask, shift, mask, counts, nblocks);
        src_k = dst_k;
        src_v = dst_v;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// ---- unique -----------------------------------------------------------------------------------------------------------
__global__ void uq_flags_kernel(const uint64_t *__restrict__ sorted, int64_t n, int64_t *__restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}
__global__ void uq_emit_kernel(const uint64_t *__restrict__ sorted, const int64_t *__restrict__ orig, const int64_t *__restrict__ excl, int64_t n,
                               int64_t *__restrict__ unq, int64_t *__restrict__ inv, int64_t *__restrict__ n_unique) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = excl[i];
        const bool head = excl[i + 1] != e;
        if (head) unq[e] = (int64_t)sorted[i];
        if (inv) inv[orig[i]] = head ? e : e - 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_unique = excl[n];
}

// ---- top-k selection ----------------------------------------------------------------------------------------------------
struct SelectState {
    uint64_t prefix;        // digits of the k-th smallest transformed key decided so far
    uint64_t decided_mask;  // which bits of prefix are decided
    int64_t k_rem;          // how many of the elements that match the prefix are still wanted
    uint32_t done_blocks;   // blocks of the current digit's launch that have added their histogram
    uint32_t pad;
    uint32_t hist[256];
};

__global__ void sel_init_kernel(SelectState *st, int64_t k) {
    if (threadIdx.x == 0) {
        st->prefix = 0ull;
        st->decided_mask = 0ull;
        st->k_rem = k;
        st->done_blocks = 0u;
    }
    st->hist[threadIdx.x] = 0u;
}
// one digit of the selection: histogram of the digit over the elements that match the prefix decided so far; the block that
// finishes last reads the histogram, picks the bucket of the k-th smallest key and extends the prefix (no second launch)
__global__ void __launch_bounds__(256) sel_digit_kernel(const uint64_t *__restrict__ vals, int64_t n, int kind, uint64_t xor_mask, int shift,
                                                       SelectState *st) {
    __shared__ uint32_t hist[256];
    __shared__ bool s_last;
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint64_t prefix = st->prefix, dmask = st->decided_mask;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = sort_key(vals[i], kind, xor_mask);
        if ((key & dmask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (hist[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], hist[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->done_blocks, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    hist[threadIdx.x] = *reinterpret_cast<volatile uint32_t *>(&st->hist[threadIdx.x]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t k = st->k_rem, before = 0;
        int b = 0;
        for (; b < 255; ++b) {
            if (before + (int64_t)hist[b] >= k) break;
            before += hist[b];
        }
        st->prefix = prefix | ((uint64_t)b << shift);
        st->decided_mask = dmask | (0xffull << shift);
        st->k_rem = k - before;
        st->done_blocks = 0u;
    }
    st->hist[threadIdx.x] = 0u;
}
// packed flags: (key < T) << 32 | (key == T); their scan gives both ranks at once
__global__ void sel_flags_kernel(const uint64_t *__restrict__ vals, int64_t n, int kind, uint64_t xor_mask, const SelectState *st,
                                 int64_t *__restrict__ flags) {
    const uint64_t T = st->prefix;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = sort_key(vals[i], kind, xor_mask);
        flags[i] = key < T ? ((int64_t)1 << 32) : (key == T ? 1 : 0);
    }
}
__global__ void sel_compact_kernel(const uint64_t *__restrict__ vals, int64_t n, int kind, uint64_t xor_mask, const SelectState *st,
                                   const int64_t *__restrict__ excl, uint64_t *__restrict__ out_bits, int64_t *__restrict__ out_idx) {
    const uint64_t T = st->prefix;
    const int64_t k_eq = st->k_rem;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t bits = vals[i], key = sort_key(bits, kind, xor_mask);
        const int64_t e = excl[i], lt = e >> 32, eq = e & 0xffffffffll;
        const bool keep = key < T || (key == T && eq < k_eq);
        if (keep) {
            const int64_t pos = lt + min(eq, k_eq);
            out_bits[pos] = bits;
            out_idx[pos] = i;
        }
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

size_t anqs_sort_workspace(int64_t n) { return n < 0 ? 0 : sort_ws_bytes(n); }

int anqs_sort_pairs_u64(const uint64_t *d_keys_in, const int64_t *d_vals_in, uint64_t *d_keys_out, int64_t *d_vals_out, int64_t n,
                        int begin_bit, int end_bit, int key_kind, uint64_t xor_mask, void *d_work, void *stream) {
    ANQS_REQUIRE(n >= 0 && n < ((int64_t)1 << 32), "element count out of range");
    ANQS_REQUIRE(0 <= begin_bit && begin_bit <= end_bit && end_bit <= 64, "bit range out of [0, 64]");
    ANQS_REQUIRE(key_kind == 0 || key_kind == 1, "key_kind must be 0 (unsigned integer) or 1 (float64)");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_keys_in && d_keys_out && d_vals_out && d_work, "null pointer");
    ANQS_REQUIRE(d_keys_in != d_keys_out && (const void *)d_vals_in != (const void *)d_vals_out, "in-place sorting is not supported");
    const int rc = sort_pairs(d_keys_in, d_vals_in, d_keys_out, d_vals_out, n, begin_bit, end_bit, key_kind, xor_mask, d_work, (cudaStream_t)stream);
    ANQS_REQUIRE(rc == 0, "kernel launch failed");
    return 0;
}

// workspace of anqs_unique_i64: sorted keys + positions + flags/scan (n + 1) + the scan's and the sort's own workspaces
static size_t unique_ws_bytes(int64_t n) {
    return 2 * align256((size_t)std::max<int64_t>(n, 1) * 8) + align256((size_t)(n + 1) * 8) + align256((size_t)(n + 1) * 8) +
           align256(anqs_scan_workspace(n)) + sort_ws_bytes(n);
}
size_t anqs_unique_workspace(int64_t n) { return n < 0 ? 0 : unique_ws_bytes(n); }

int anqs_unique_i64(const int64_t *d_in, int64_t n, int end_bit, int64_t *d_unq, int64_t *d_inv, int64_t *d_n_unique, void *d_work,
                    void *stream) {
    ANQS_REQUIRE(n >= 0 && n < ((int64_t)1 << 32), "element count out of range");
    ANQS_REQUIRE(end_bit >= 0 && end_bit <= 64, "end_bit out of [0, 64]");
    ANQS_REQUIRE(d_n_unique, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        ANQS_CUDA(cudaMemsetAsync(d_n_unique, 0, sizeof(int64_t), s));
        return 0;
    }
    ANQS_REQUIRE(d_in && d_unq && d_work, "null pointer");
    unsigned char *w = (unsigned char *)d_work;
    uint64_t *sorted = (uint64_t *)w;
    w += align256((size_t)n * 8);
    int64_t *orig = (int64_t *)w;
    w += align256((size_t)n * 8);
    int64_t *flags = (int64_t *)w;
    w += align256((size_t)(n + 1) * 8);
    int64_t *excl = (int64_t *)w;
    w += align256((size_t)(n + 1) * 8);
    void *scan_ws = w;
    w += align256(anqs_scan_workspace(n));
    // signed ascending order (what torch.unique returns, HS:215-228): unsigned order of key ^ 2^63 when bit 63 takes part
    const uint64_t xor_mask = end_bit == 64 ? (1ull << 63) : 0ull;
    ANQS_REQUIRE(sort_pairs((const uint64_t *)d_in, nullptr, sorted, orig, n, 0, end_bit, 0, xor_mask, w, s) == 0, "kernel launch failed");
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    uq_flags_kernel<<<grid, 256, 0, s>>>(sorted, n, flags);
    ANQS_LAUNCH_CHECK();
    if (anqs_exclusive_scan_i64(flags, excl, n, scan_ws, stream) != 0) return 2;
    uq_emit_kernel<<<grid, 256, 0, s>>>(sorted, orig, excl, n, d_unq, d_inv, d_n_unique);
    ANQS_LAUNCH_CHECK();
    return 0;
}

// workspace of anqs_topk_f64: selection state + flags/scan (n + 1) + survivors (k pairs) + scan and sort workspaces
static size_t topk_ws_bytes(int64_t n, int64_t k) {
    return align256(sizeof(SelectState)) + 2 * align256((size_t)(n + 1) * 8) + 2 * align256((size_t)std::max<int64_t>(k, 1) * 8) +
           align256(anqs_scan_workspace(n)) + sort_ws_bytes(std::max(n, k));
}
size_t anqs_topk_workspace(int64_t n, int64_t k) { return (n < 0 || k < 0) ? 0 : topk_ws_bytes(n, k); }

int anqs_topk_f64(const double *d_vals, int64_t n, int64_t k, int sorted, double *d_top_vals, int64_t *d_top_idx, void *d_work, void *stream) {
    ANQS_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "element count out of range");
    ANQS_REQUIRE(k >= 0 && k <= n, "need 0 <= k <= n");
    if (k == 0) return 0;
    ANQS_REQUIRE(d_vals && d_top_vals && d_top_idx && d_work, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t desc = ~0ull;  // descending values = ascending ~key
    unsigned char *w = (unsigned char *)d_work;
    SelectState *st = (SelectState *)w;
    w += align256(sizeof(SelectState));
    int64_t *flags = (int64_t *)w;
    w += align256((size_t)(n + 1) * 8);
    int64_t *excl = (int64_t *)w;
    w += align256((size_t)(n + 1) * 8);
    uint64_t *kept_bits = (uint64_t *)w;
    w += align256((size_t)k * 8);
    int64_t *kept_idx = (int64_t *)w;
    w += align256((size_t)k * 8);
    void *scan_ws = w;
    w += align256(anqs_scan_workspace(n));
    const uint64_t *bits = (const uint64_t *)d_vals;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    if (k == n && !sorted) {  // everything is kept and no order is asked for
        copy_pairs_kernel<<<grid, 256, 0, s>>>(bits, nullptr, (uint64_t *)d_top_vals, d_top_idx, n);
        ANQS_LAUNCH_CHECK();
        return 0;
    }
    if (sorted && (n <= 2 * k || n <= RS_SMALL_MAX)) {  // nothing to gain from selecting first (small inputs sort in one launch)
        ANQS_REQUIRE(sort_pairs(bits, nullptr, (uint64_t *)flags, excl, n, 0, 64, 1, desc, w, s) == 0, "kernel launch failed");
        ANQS_CUDA(cudaMemcpyAsync(d_top_vals, flags, (size_t)k * 8, cudaMemcpyDeviceToDevice, s));
        ANQS_CUDA(cudaMemcpyAsync(d_top_idx, excl, (size_t)k * 8, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    sel_init_kernel<<<1, 256, 0, s>>>(st, k);
    for (int shift = 56; shift >= 0; shift -= 8) sel_digit_kernel<<<grid, 256, 0, s>>>(bits, n, 1, desc, shift, st);
    sel_flags_kernel<<<grid, 256, 0, s>>>(bits, n, 1, desc, st, flags);
    ANQS_LAUNCH_CHECK();
    if (anqs_exclusive_scan_i64(flags, excl, n, scan_ws, stream) != 0) return 2;
    if (!sorted) {  // survivors in position order: a valid top-k set, one sort cheaper
        sel_compact_kernel<<<grid, 256, 0, s>>>(bits, n, 1, desc, st, excl, (uint64_t *)d_top_vals, d_top_idx);
        ANQS_LAUNCH_CHECK();
        return 0;
    }
    sel_compact_kernel<<<grid, 256, 0, s>>>(bits, n, 1, desc, st, excl, kept_bits, kept_idx);
    ANQS_LAUNCH_CHECK();
    ANQS_REQUIRE(sort_pairs(kept_bits, kept_idx, (uint64_t *)d_top_vals, d_top_idx, k, 0, 64, 1, desc, w, s) == 0, "kernel launch failed");
    return 0;
}

}  // extern "C"
