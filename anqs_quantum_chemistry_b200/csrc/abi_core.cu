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


// ---- product layout of the XY masks (Tables::prod_*, consumed by k1_fused.cu) --------------------------------
constexpr uint32_t PROD_TILE_MAX = 200 * 1024;  // bytes of shared memory one tile may take
constexpr uint32_t PROD_ROW_CHUNK = 2048;       // longer rows are split so that tiles pack evenly
constexpr uint32_t PROD_SINGLE_MAX = 2;         // rows with <= this many members become singleton records

struct HostProd {
    std::vector<ProdTile> tiles;
    std::vector<uint8_t> blob;
    std::vector<uint32_t> row_u, mem_u;
    uint32_t tile_bytes_max = 0;
};

static void build_product_layout(const std::vector<uint2> &mab, int64_t U, HostProd &out) {
    // masks sorted by (alpha part, spread bits of the member hash, beta part)
    struct M { uint32_t pa, mb, hash, u; };
    std::vector<M> ms((size_t)U);
    for (int64_t u = 0; u < U; ++u) {
        uint32_t pa = mab[u].x, mb = mab[u].y;
        ms[u] = {pa, mb, (lin_host(LIN_POSA, pa) & POSA_MASK) ^ (lin_host(LIN_POSB, mb) & POSB_MASK), (uint32_t)u};
    }
    std::sort(ms.begin(), ms.end(), [](const M &x, const M &y) {
        if (x.pa != y.pa) return x.pa < y.pa;
        uint32_t sx = x.hash >> 15, sy = y.hash >> 15;
        if (sx != sy) return sx < sy;
        return x.mb < y.mb;
    });
    // row chunks: [begin, end) ranges of ms with one alpha part and at most PROD_ROW_CHUNK members
    struct Chunk { size_t begin, end; };
    std::vector<Chunk> multi, single;
    for (size_t i = 0; i < ms.size();) {
        size_t j = i;
        while (j < ms.size() && ms[j].pa == ms[i].pa) ++j;
        if (j - i <= PROD_SINGLE_MAX) {
            for (size_t k = i; k < j; ++k) single.push_back({k, k + 1});
        } else {
            for (size_t k = i; k < j; k += PROD_ROW_CHUNK) multi.push_back({k, std::min(j, k + (size_t)PROD_ROW_CHUNK)});
        }
        i = j;
    }
    auto chunk_bytes = [](const Chunk &c, bool is_multi) -> size_t {
        return sizeof(RowRec) + (is_multi ? (c.end - c.begin) * sizeof(MemRec) : 0);
    };
    size_t total = 0;
    for (auto &c : multi) total += chunk_bytes(c, true);
    for (auto &c : single) total += chunk_bytes(c, false);
    const size_t n_tiles = std::max<size_t>(1, (total + PROD_TILE_MAX - 1) / PROD_TILE_MAX);
    const size_t target = (total + n_tiles - 1) / n_tiles;
    // greedy packing in order (multi rows first, then singletons); a tile closes once it reaches the target
    size_t im = 0, is = 0;
    uint32_t row_base = 0, member_base = 0;
    while (im < multi.size() || is < single.size()) {
        std::vector<Chunk> tm, ts;
        size_t bytes = 0;
        while (im < multi.size() && bytes < target && bytes + chunk_bytes(multi[im], true) <= PROD_TILE_MAX) {
            bytes += chunk_bytes(multi[im], true);
            tm.push_back(multi[im++]);
        }
        if (im == multi.size()) {
            while (is < single.size() && bytes + sizeof(RowRec) <= PROD_TILE_MAX && bytes < target) {
                bytes += sizeof(RowRec);
                ts.push_back(single[is++]);
            }
        }
        ProdTile tile{};
        tile.n_multi = (uint32_t)tm.size();
        tile.n_single = (uint32_t)ts.size();
        tile.row_base = row_base;
        tile.member_base = member_base;
        std::vector<RowRec> rows;
        std::vector<MemRec> mems;
        for (auto &c : tm) {
            const uint32_t pa = ms[c.begin].pa;
            rows.push_back({pa, lin_host(LIN_LINE, pa), (uint32_t)mems.size(), (uint32_t)(c.end - c.begin)});
            out.row_u.push_back(0);
            for (size_t k = c.begin; k < c.end; ++k) {
                mems.push_back({ms[k].mb, ms[k].hash});
                out.mem_u.push_back(ms[k].u);
            }
        }
        for (auto &c : ts) {
            const M &m = ms[c.begin];
            rows.push_back({m.pa, lin_host(LIN_LINE, m.pa), m.mb, m.hash});
            out.row_u.push_back(m.u);
        }
        tile.n_members = (uint32_t)mems.size();
        size_t off = (out.blob.size() + 127) / 128 * 128;
        size_t nbytes = rows.size() * sizeof(RowRec) + mems.size() * sizeof(MemRec);
        nbytes = (nbytes + 15) / 16 * 16;
        out.blob.resize(off + nbytes, 0);
        std::memcpy(out.blob.data() + off, rows.data(), rows.size() * sizeof(RowRec));
        std::memcpy(out.blob.data() + off + rows.size() * sizeof(RowRec), mems.data(), mems.size() * sizeof(MemRec));
        tile.blob_off = (uint32_t)off;
        tile.blob_bytes = (uint32_t)nbytes;
        out.tile_bytes_max = std::max(out.tile_bytes_max, tile.blob_bytes);
        out.tiles.push_back(tile);
        row_base += (uint32_t)rows.size();
        member_base += (uint32_t)mems.size();
    }
    if (out.mem_u.empty()) out.mem_u.push_back(0);
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
    HostProd prod;
    build_product_layout(mab, U, prod);
    t->n_tiles = (int)prod.tiles.size();
    t->tile_bytes_max = (int)prod.tile_bytes_max;
    if (e == cudaSuccess) e = up((void **)&t->prod_tiles, prod.tiles.data(), prod.tiles.size() * sizeof(ProdTile));
    if (e == cudaSuccess) e = up((void **)&t->prod_blob, prod.blob.data(), prod.blob.size());
    if (e == cudaSuccess) e = up((void **)&t->prod_row_u, prod.row_u.data(), prod.row_u.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = up((void **)&t->prod_mem_u, prod.mem_u.data(), prod.mem_u.size() * sizeof(uint32_t));
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
    cudaFree(t->prod_tiles);
    cudaFree(t->prod_blob);
    cudaFree(t->prod_row_u);
    cudaFree(t->prod_mem_u);
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
