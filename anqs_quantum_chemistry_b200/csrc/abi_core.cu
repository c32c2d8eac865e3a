// libanqs_b200.so -- error handling, device check, Hamiltonian table handle, popcount, scan.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace anqs {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int sm_count_of_current_device() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- A4 popcount: one 16-byte load + two POPC pairs + one 16-byte store per thread iteration -------
__global__ void popcount_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n2 = n >> 1;
    const longlong2 *in2 = reinterpret_cast<const longlong2 *>(in);
    longlong2 *out2 = reinterpret_cast<longlong2 *>(out);
    for (int64_t k = i; k < n2; k += stride) {
        longlong2 v = in2[k];
        v.x = __popcll((unsigned long long)v.x);
        v.y = __popcll((unsigned long long)v.y);
        out2[k] = v;
    }
    if ((n & 1) && i == 0) out[n - 1] = __popcll((unsigned long long)in[n - 1]);
}

__global__ void popcount_kernel_unaligned(const int64_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        out[k] = __popcll((unsigned long long)in[k]);
}

// ---- exclusive scan (three phases; n is at most a few 1e7 so one block scans the block sums) -------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *warp_sums) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t s = lane < (SCAN_BLOCK / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        if (lane < (SCAN_BLOCK / 32)) warp_sums[lane] = s;
    }
    __syncthreads();
    int64_t base = warp > 0 ? warp_sums[warp - 1] : 0;
    *total = warp_sums[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void scan_tile_sums(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int64_t total;
    block_exclusive_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_of_tile_sums(int64_t *tile_sums, int64_t ntiles) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t carry = 0;
    for (int64_t lo = 0; lo < ntiles; lo += SCAN_BLOCK) {
        int64_t i = lo + threadIdx.x;
        int64_t v = i < ntiles ? tile_sums[i] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, &total, warp_sums);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void scan_apply(const int64_t *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_offsets,
                           int64_t *__restrict__ out) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, &total, warp_sums) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_abi_version(void) { return ANQS_ABI_VERSION; }

const char *anqs_last_error(void) { return g_last_error.c_str(); }

int anqs_device_check(int device, int *sm_count, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    ANQS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    ANQS_REQUIRE(prop.major == 10, "libanqs_b200 is built for sm_100a only; device is compute capability " +
                                       std::to_string(prop.major) + "." + std::to_string(prop.minor));
    return 0;
}

int anqs_popcount_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative element count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_in && d_out, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = sm_count_of_current_device();
    bool aligned = (((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0;
    int64_t work = aligned ? (n + 1) / 2 : n;
    int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)sms * 16);
    if (aligned)
        popcount_kernel<<<blocks, 256, 0, s>>>(d_in, d_out, n);
    else
        popcount_kernel_unaligned<<<blocks, 256, 0, s>>>(d_in, d_out, n);
    ANQS_LAUNCH_CHECK();
    return 0;
}

size_t anqs_scan_workspace(int64_t n) {
    int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(ntiles + 1) * sizeof(int64_t);
}

int anqs_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *d_work, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative element count");
    ANQS_REQUIRE(d_out, "null output");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        ANQS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), s));
        return 0;
    }
    ANQS_REQUIRE(d_in && d_work, "null pointer");
    int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t *tile_sums = (int64_t *)d_work;
    scan_tile_sums<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(d_in, n, tile_sums);
    ANQS_LAUNCH_CHECK();
    scan_of_tile_sums<<<1, SCAN_BLOCK, 0, s>>>(tile_sums, ntiles);
    ANQS_LAUNCH_CHECK();
    scan_apply<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(d_in, n, tile_sums, d_out);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_tables_create(anqs_tables_t **out, int qubit_num, int64_t U, int64_t T, const int64_t *h_unq_xy,
                       const int64_t *h_yz_num, const int64_t *h_yz_start, const int64_t *h_yz,
                       const double *h_weights) {
    ANQS_REQUIRE(out, "null handle pointer");
    ANQS_REQUIRE(qubit_num >= 1 && qubit_num <= 64, "qubit_num must be in [1, 64] (single-word indices)");
    ANQS_REQUIRE(U >= 1 && T >= U, "need 1 <= U <= T");
    ANQS_REQUIRE(T < (int64_t)1 << 31, "too many terms");
    ANQS_REQUIRE(h_unq_xy && h_yz_num && h_yz_start && h_yz && h_weights, "null table pointer");
    Tables *t = new Tables();
    std::memset(t, 0, sizeof(Tables));
    t->qubit_num = qubit_num;
    t->U = U;
    t->T = T;
    t->U_pad = (U + 1023) / 1024 * 1024;
    t->row_words = t->U_pad / 32;
    ANQS_CUDA(cudaGetDevice(&t->device));

    std::vector<uint64_t> xy(t->U_pad, 0);
    std::vector<uint2> mab(t->U_pad, make_uint2(0, 0));
    std::vector<int2> grp(t->U_pad, make_int2(0, 0));
    int max_group = 0;
    for (int64_t u = 0; u < U; ++u) {
        uint64_t m = (uint64_t)h_unq_xy[u];
        if (u > 0 && !(h_unq_xy[u - 1] < h_unq_xy[u])) {
            delete t;
            ANQS_REQUIRE(false, "unq_xy_masks must be strictly ascending (signed order)");
        }
        if (h_yz_start[u] < 0 || h_yz_num[u] < 0 || h_yz_start[u] + h_yz_num[u] > T) {
            delete t;
            ANQS_REQUIRE(false, "YZ group out of range");
        }
        xy[u] = m;
        mab[u] = make_uint2(compress_even_bits(m), compress_even_bits(m >> 1));
        grp[u] = make_int2((int)h_yz_start[u], (int)h_yz_num[u]);
        max_group = std::max(max_group, (int)h_yz_num[u]);
    }
    t->max_group = max_group;
    bool real = true;
    std::vector<double> wre(T), wim(T);
    std::vector<ulonglong2> rec(T);
    std::vector<uint64_t> yzd(T);
    for (int64_t k = 0; k < T; ++k) {
        wre[k] = h_weights[2 * k];
        wim[k] = h_weights[2 * k + 1];
        if (wim[k] != 0.0) real = false;
        unsigned long long bits;
        std::memcpy(&bits, &wre[k], 8);
        yzd[k] = deinterleave((uint64_t)h_yz[k]);
        rec[k] = make_ulonglong2((unsigned long long)yzd[k], bits);
    }
    t->weights_real = real ? 1 : 0;

    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = up((void **)&t->xy, xy.data(), xy.size() * sizeof(uint64_t));
    if (e == cudaSuccess) e = up((void **)&t->mab, mab.data(), mab.size() * sizeof(uint2));
    if (e == cudaSuccess) e = up((void **)&t->grp, grp.data(), grp.size() * sizeof(int2));
    if (e == cudaSuccess) e = up((void **)&t->yz_d, yzd.data(), (size_t)T * sizeof(uint64_t));
    if (e == cudaSuccess) e = up((void **)&t->w_re, wre.data(), (size_t)T * sizeof(double));
    if (e == cudaSuccess && !real) e = up((void **)&t->w_im, wim.data(), (size_t)T * sizeof(double));
    if (e == cudaSuccess && real) e = up((void **)&t->term_real, rec.data(), (size_t)T * sizeof(ulonglong2));
    if (e != cudaSuccess) {
        anqs_tables_destroy((anqs_tables_t *)t);
        set_error(std::string("anqs_tables_create: device upload failed: ") + cudaGetErrorString(e));
        return 2;
    }
    *out = (anqs_tables_t *)t;
    return 0;
}

int anqs_tables_destroy(anqs_tables_t *h) {
    if (!h) return 0;
    Tables *t = (Tables *)h;
    cudaFree(t->xy);
    cudaFree(t->mab);
    cudaFree(t->grp);
    cudaFree(t->yz_d);
    cudaFree(t->w_re);
    cudaFree(t->w_im);
    cudaFree(t->term_real);
    delete t;
    return 0;
}

int anqs_tables_info(const anqs_tables_t *h, int *qubit_num, int64_t *U, int64_t *T, int *weights_real,
                     int64_t *bitmap_row_words) {
    ANQS_REQUIRE(h, "null handle");
    const Tables *t = (const Tables *)h;
    if (qubit_num) *qubit_num = t->qubit_num;
    if (U) *U = t->U;
    if (T) *T = t->T;
    if (weights_real) *weights_real = t->weights_real;
    if (bitmap_row_words) *bitmap_row_words = t->row_words;
    return 0;
}

}  // extern "C"
