// Kernel family 3, backward: the reductions over the batch that turn the per-sample chain (k3_made_bwd.cu) into parameter
// gradients (reference: autograd through MLP:217-246 — torch's addmm backward, `grad W = dY^T h`, `grad b = sum_s dY`).
//
// Every one of them is C[M x N] (+)= A^T B with A [K x M] and B [K x N] both stored sample-major (one row per sample), K = the
// batch (10^5..10^6), M <= 640, N <= 64: a tiny output and an enormous reduction dimension.  Library fp64 GEMMs pick
// 32x32 tiles for that shape (measured 5 TFLOP/s on B200); here the output is cut into 128 x 64 tiles, the batch into S
// slices so that all CTAs together are one wave (a 128-row tile gets 3/2 of the slices of a 64-row tile, which runs 4 x 8
// per thread), and each CTA keeps its whole output tile in registers (mma accumulator fragments) while its slice of A and B
// streams through a double-buffered shared-memory tile by cp.async.  The column sums of A (bias gradients) ride along
// from the same shared-memory tile.  Partial tiles go to a caller-supplied workspace and a second small kernel adds them up
// in a fixed order: the result is deterministic.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "../../include/anqs_b200.h"

namespace anqs {

constexpr int BR_MT = 128, BR_NT = 64, BR_KT = 16, BR_THREADS = 128, BR_MAX_TILES = 64;
constexpr int BR_TILE_DOUBLES = BR_MT * BR_NT + BR_MT;   // partial C tile + partial column sums

struct BrTile {
    const double *A, *B;
    double *C, *colsum;
    int lda, ldb, ldc, M, N;   // M <= BR_MT rows of C (= columns of A) in this tile, N <= BR_NT
};
struct BrBatch {
    BrTile t[BR_MAX_TILES];
    int first[BR_MAX_TILES + 1];   // CTAs first[i] .. first[i + 1] - 1 are the K slices of tile i
};

// 8-byte asynchronous copy global -> shared, zero-filled when !ok (src-size 0): no registers held across the compute phase
__device__ __forceinline__ void br_cp_async8(double *dst, const double *src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// One K slice of one output tile with MT (128 or 64) rows.  The products run on the FP64 tensor-core path (mma.sync m8n8k4 f64:
// the DFMA rate at an eighth of the issue slots, scripts/microbench_dmma.cu): warp w of the four owns rows w MT/4 ..
// (w + 1) MT/4 - 1 of the tile (MT / 32 row fragments) and all 64 columns (8 column fragments), MT / 32 + 8 shared-memory
// loads of 8 bytes per 8 MT / 32 mma.  Row k of a shared tile is stored rotated by 4 (k mod 4) columns, so that the
// half-warp's fragment addresses (k = lane % 4, column = lane / 4) fall on 16 different bank pairs without padding.
__device__ __forceinline__ void br_mma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MT>
__device__ __forceinline__ void br_slice(const BrTile &T, int64_t k0, int64_t k1, double (*As)[BR_KT][BR_MT], double (*Bs)[BR_KT][BR_NT],
                                         double *__restrict__ out) {
    constexpr int RM = MT / 32;              // row fragments (8 rows each) per warp
    const int lda = T.lda, ldb = T.ldb, M = T.M, N = T.N;
    const bool want_sums = T.colsum != nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;                 // A fragment: (row fr, k fk); B fragment: (k fk, column fr)
    const int b_col = tid & (BR_NT - 1), b_row = tid >> 6;   // B tile: 8 elements per thread, rows b_row + 2 j
    const bool a_ok = tid < M, b_ok = b_col < N;             // A tile: 16 elements per thread, column tid, rows j
    const double *Ap = T.A + (a_ok ? tid : 0), *Bp = T.B + (b_ok ? b_col : 0);

    double acc[RM][8][2];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[i][t][0] = acc[i][t][1] = 0.0;
    double csum = 0.0;

    auto fetch = [&](int64_t kb, int buf) {
        if (MT == 128 || tid < 64) {
#pragma unroll
            for (int j = 0; j < BR_KT; ++j) {
                const int64_t k = kb + j;
                const bool ok = a_ok && k < k1;
                br_cp_async8(&As[buf][j][(tid + 4 * (j & 3)) & (BR_MT - 1)], Ap + (ok ? k : k0) * lda, ok);
            }
        }
#pragma unroll
        for (int j = 0; j < BR_KT / 2; ++j) {
            const int r = b_row + 2 * j;
            const int64_t k = kb + r;
            const bool ok = b_ok && k < k1;
            br_cp_async8(&Bs[buf][r][(b_col + 4 * (r & 3)) & (BR_NT - 1)], Bp + (ok ? k : k0) * ldb, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    // this lane's fragment columns (rotation of its k row included)
    int a_col[RM], b_colf[8];
#pragma unroll
    for (int i = 0; i < RM; ++i) a_col[i] = (warp * 8 * RM + 8 * i + fr + 4 * fk) & (BR_MT - 1);
#pragma unroll
    for (int t = 0; t < 8; ++t) b_colf[t] = (8 * t + fr + 4 * fk) & (BR_NT - 1);

    if (k0 < k1) fetch(k0, 0);
    int buf = 0;
    for (int64_t kb = k0; kb < k1; kb += BR_KT, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                                  // tile `buf` has landed; everyone is done with tile `buf ^ 1`
        if (kb + BR_KT < k1) fetch(kb + BR_KT, buf ^ 1);
#pragma unroll
        for (int kk = 0; kk < BR_KT; kk += 4) {
            double a[RM], b[8];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[buf][kk + fk][a_col[i]];
#pragma unroll
            for (int t = 0; t < 8; ++t) b[t] = Bs[buf][kk + fk][b_colf[t]];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int t = 0; t < 8; ++t) br_mma(acc[i][t], a[i], b[t]);
        }
        if (want_sums && tid < MT) {
#pragma unroll
            for (int kk = 0; kk < BR_KT; ++kk) csum += As[buf][kk][(tid + 4 * (kk & 3)) & (BR_MT - 1)];
        }
    }

#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int r = warp * 8 * RM + 8 * i + fr;
#pragma unroll
        for (int t = 0; t < 8; ++t)
            *reinterpret_cast<double2 *>(out + r * BR_NT + 8 * t + 2 * fk) = make_double2(acc[i][t][0], acc[i][t][1]);
    }
    if (tid < MT) out[BR_MT * BR_NT + tid] = csum;
}

// MT_MAX = 128: tiles of either height in one launch, two CTAs per SM; MT_MAX = 64: every tile has <= 64 rows, three CTAs per SM.
template <int MT_MAX>
__global__ void __launch_bounds__(BR_THREADS, MT_MAX == 128 ? 2 : 3)
batch_reduce_gemm_kernel(const BrBatch P, int n_tiles, int64_t K, double *__restrict__ ws) {
    __shared__ __align__(16) double As[2][BR_KT][BR_MT];
    __shared__ __align__(16) double Bs[2][BR_KT][BR_NT];
    int tile = 0;
    while (tile + 1 < n_tiles && (int)blockIdx.x >= P.first[tile + 1]) ++tile;
    const int S = P.first[tile + 1] - P.first[tile], split = (int)blockIdx.x - P.first[tile];
    const int64_t kps = ((K + S - 1) / S + BR_KT - 1) / BR_KT * BR_KT;
    const int64_t k0 = split * kps, k1 = min(K, k0 + kps);
    double *out = ws + (size_t)blockIdx.x * BR_TILE_DOUBLES;
    if (MT_MAX == 64 || P.t[tile].M <= 64)
        br_slice<64>(P.t[tile], k0, k1, As, Bs, out);
    else
        br_slice<128>(P.t[tile], k0, k1, As, Bs, out);
}

// C (+)= sum over the partial tiles of the K slices, in slice order; eight CTAs per tile.
__global__ void __launch_bounds__(256)
batch_reduce_finish_kernel(const BrBatch P, int accumulate, const double *__restrict__ ws) {
    const BrTile T = P.t[blockIdx.y];
    const int S = P.first[blockIdx.y + 1] - P.first[blockIdx.y];
    const double *base = ws + (size_t)P.first[blockIdx.y] * BR_TILE_DOUBLES;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < BR_TILE_DOUBLES; e += gridDim.x * 256) {
        const bool is_sum = e >= BR_MT * BR_NT;
        const int r = is_sum ? e - BR_MT * BR_NT : e / BR_NT, c = is_sum ? 0 : e % BR_NT;
        if (r >= T.M || c >= T.N || (is_sum && T.colsum == nullptr)) continue;
        double s = 0.0;
        for (int k = 0; k < S; ++k) s += base[(size_t)k * BR_TILE_DOUBLES + e];
        double *dst = is_sum ? T.colsum + r : T.C + (size_t)r * T.ldc + c;
        *dst = accumulate ? *dst + s : s;
    }
}

// Cuts the problems into tiles and the tiles into launches of <= BR_MAX_TILES; every launch is one wave: a 128-row tile
// gets 3/2 of the K slices of a 64-row tile, so that all CTAs of the launch take about the same time.
struct BrLaunch {
    BrBatch b;
    int n_tiles, n_ctas, all_half;
};
static int br_plan(const anqs_brg_problem_t *p, int n, int64_t K, std::vector<BrLaunch> &launches) {
    std::vector<BrTile> tiles;
    for (int i = 0; i < n; ++i) {
        if (!(p[i].A && p[i].B && p[i].C && p[i].M > 0 && p[i].N > 0 && p[i].N <= BR_NT && p[i].lda >= p[i].M && p[i].ldb >= p[i].N &&
              p[i].ldc >= p[i].N))
            return 1;
        for (int m0 = 0; m0 < p[i].M; m0 += BR_MT) {
            BrTile t;
            t.A = p[i].A + m0;
            t.B = p[i].B;
            t.C = p[i].C + (size_t)m0 * p[i].ldc;
            t.colsum = p[i].colsum ? p[i].colsum + m0 : nullptr;
            t.lda = p[i].lda, t.ldb = p[i].ldb, t.ldc = p[i].ldc;
            t.M = std::min(BR_MT, p[i].M - m0), t.N = p[i].N;
            tiles.push_back(t);
        }
    }
    const int sms = sm_count_of_current_device();
    const int64_t most = std::max<int64_t>(1, K / (4 * BR_KT));   // at least four K tiles per CTA
    for (size_t t0 = 0; t0 < tiles.size(); t0 += BR_MAX_TILES) {
        BrLaunch L;
        L.n_tiles = (int)std::min<size_t>(BR_MAX_TILES, tiles.size() - t0);
        int units = 0, n_half = 0;   // measured cost of a K row: 64-row tile : 128-row tile = 2 : 3 (same fixed work per K tile)
        for (int i = 0; i < L.n_tiles; ++i) {
            units += tiles[t0 + i].M <= 64 ? 2 : 3;
            n_half += tiles[t0 + i].M <= 64;
        }
        L.all_half = n_half == L.n_tiles;
        const int64_t slots = (int64_t)sms * (L.all_half ? 3 : 2);
        L.b.first[0] = 0;
        for (int i = 0; i < L.n_tiles; ++i) {
            L.b.t[i] = tiles[t0 + i];
            const int64_t s = std::max<int64_t>(1, std::min(most, (tiles[t0 + i].M <= 64 ? 2 : 3) * slots / units));
            L.b.first[i + 1] = L.b.first[i] + (int)s;
        }
        L.n_ctas = L.b.first[L.n_tiles];
        launches.push_back(L);
    }
    return 0;
}
static int64_t br_workspace(const std::vector<BrLaunch> &launches) {
    int64_t ctas = 0;
    for (const BrLaunch &L : launches) ctas += L.n_ctas;
    return ctas * BR_TILE_DOUBLES * (int64_t)sizeof(double);
}

}  // namespace anqs

using namespace anqs;

extern "C" int64_t anqs_batch_reduce_workspace(const anqs_brg_problem_t *problems, int n_problems, int64_t K) {
    if (!problems || n_problems <= 0 || K <= 0) return 0;
    std::vector<BrLaunch> launches;
    if (br_plan(problems, n_problems, K, launches)) return -1;
    return br_workspace(launches);
}

extern "C" int anqs_batch_reduce_gemm(const anqs_brg_problem_t *problems, int n_problems, int64_t K, int accumulate, void *workspace,
                                      int64_t workspace_bytes, void *stream) {
    ANQS_REQUIRE(problems && n_problems > 0, "no problems");
    ANQS_REQUIRE(K > 0, "empty batch");
    std::vector<BrLaunch> launches;
    ANQS_REQUIRE(br_plan(problems, n_problems, K, launches) == 0, "bad problem (null pointer, N > 64 or leading dimension too small)");
    ANQS_REQUIRE(workspace && workspace_bytes >= br_workspace(launches), "workspace too small");
    double *ws = (double *)workspace;
    for (const BrLaunch &L : launches) {
        if (L.all_half)
            batch_reduce_gemm_kernel<64><<<L.n_ctas, BR_THREADS, 0, (cudaStream_t)stream>>>(L.b, L.n_tiles, K, ws);
        else
            batch_reduce_gemm_kernel<128><<<L.n_ctas, BR_THREADS, 0, (cudaStream_t)stream>>>(L.b, L.n_tiles, K, ws);
        batch_reduce_finish_kernel<<<dim3(8, L.n_tiles), 256, 0, (cudaStream_t)stream>>>(L.b, accumulate, ws);
        ws += (size_t)L.n_ctas * BR_TILE_DOUBLES;
    }
    ANQS_CUDA(cudaGetLastError());
    return 0;
}
