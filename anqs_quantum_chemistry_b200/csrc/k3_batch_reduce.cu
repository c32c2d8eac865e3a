// Kernel family 3, backward: the reductions over the batch that turn the per-sample chain (k3_made_bwd.cu) into parameter
// gradients (reference: autograd through MLP:217-246 — torch's addmm backward, `grad W = dY^T h`, `grad b = sum_s dY`).
//
// Every one of them is C[M x N] (+)= A^T B with A [K x M] and B [K x N] both stored sample-major (one row per sample), K = the
// batch (10^5..10^6), M <= 640, N <= 64: a tiny output and an enormous reduction dimension.  Library fp64 GEMMs pick
// 32x32 tiles for that shape (measured 5 TFLOP/s on B200); here the output is cut into 128 x 64 tiles, the batch into S
// slices so that tiles x S is one wave of two CTAs per SM, and each CTA keeps its whole output tile in registers (8 x 8 per
// thread) while its slice of A and B streams through a double-buffered shared-memory tile by cp.async — 8 LDS.128 per
// 64 DFMA, so the FP64 pipe and not the shared-memory pipe is the limit.  The column sums of A (bias gradients) ride along
// from the same shared-memory tile.  Partial tiles go to a caller-supplied workspace and a second small kernel adds them up
// in a fixed order: the result is deterministic.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "../../include/anqs_b200.h"

namespace anqs {

constexpr int BR_MT = 128, BR_NT = 64, BR_KT = 16, BR_THREADS = 128, BR_MAX_TILES = 32;
constexpr int BR_TILE_DOUBLES = BR_MT * BR_NT + BR_MT;   // partial C tile + partial column sums

struct BrTile {
    const double *A, *B;
    double *C, *colsum;
    int lda, ldb, ldc, M, N;   // M <= BR_MT rows of C (= columns of A) in this tile, N <= BR_NT
};
struct BrBatch {
    BrTile t[BR_MAX_TILES];
};

// 8-byte asynchronous copy global -> shared, zero-filled when !ok (src-size 0): no registers held across the compute phase
__device__ __forceinline__ void br_cp_async8(double *dst, const double *src, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(BR_THREADS, 2)
batch_reduce_gemm_kernel(const BrBatch P, int64_t K, int64_t k_per_split, double *__restrict__ ws) {
    __shared__ __align__(16) double As[2][BR_KT][BR_MT];
    __shared__ __align__(16) double Bs[2][BR_KT][BR_NT];
    const int tile = blockIdx.y;
    const double *const tA = P.t[tile].A, *const tB = P.t[tile].B;
    const int lda = P.t[tile].lda, ldb = P.t[tile].ldb, M = P.t[tile].M, N = P.t[tile].N;
    const bool want_sums = P.t[tile].colsum != nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ig = lane & 7;                 // 8 column groups: columns 2 ig + {0,1} + 16 j, j < 4
    const int og = warp * 4 + (lane >> 3);   // 16 row groups:   rows    2 og + {0,1} + 32 j, j < 4
    const int64_t k0 = (int64_t)blockIdx.x * k_per_split, k1 = min(K, k0 + k_per_split);
    const int b_col = tid & (BR_NT - 1), b_row = tid >> 6;   // B tile: 8 elements per thread, rows b_row + 2 j
    const bool a_ok = tid < M, b_ok = b_col < N;             // A tile: 16 elements per thread, column tid, rows j
    const double *Ap = tA + (a_ok ? tid : 0), *Bp = tB + (b_ok ? b_col : 0);

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    double csum = 0.0;

    auto fetch = [&](int64_t kb, int buf) {
#pragma unroll
        for (int j = 0; j < BR_KT; ++j) {
            const int64_t k = kb + j;
            const bool ok = a_ok && k < k1;
            br_cp_async8(&As[buf][j][tid], Ap + (ok ? k : k0) * lda, ok);
        }
#pragma unroll
        for (int j = 0; j < BR_KT / 2; ++j) {
            const int64_t k = kb + b_row + 2 * j;
            const bool ok = b_ok && k < k1;
            br_cp_async8(&Bs[buf][b_row + 2 * j][b_col], Bp + (ok ? k : k0) * ldb, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    if (k0 < k1) fetch(k0, 0);
    int buf = 0;
    for (int64_t kb = k0; kb < k1; kb += BR_KT, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                                  // tile `buf` has landed; everyone is done with tile `buf ^ 1`
        if (kb + BR_KT < k1) fetch(kb + BR_KT, buf ^ 1);
#pragma unroll
        for (int kk = 0; kk < BR_KT; ++kk) {
            double a[8], b[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(&As[buf][kk][2 * og + 32 * j]);
                a[2 * j] = v.x;
                a[2 * j + 1] = v.y;
                const double2 w = *reinterpret_cast<const double2 *>(&Bs[buf][kk][2 * ig + 16 * j]);
                b[2 * j] = w.x;
                b[2 * j + 1] = w.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (want_sums) {
#pragma unroll
            for (int kk = 0; kk < BR_KT; ++kk) csum += As[buf][kk][tid];
        }
    }

    double *out = ws + ((size_t)tile * gridDim.x + blockIdx.x) * BR_TILE_DOUBLES;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 2 * og + (i & 1) + 32 * (i >> 1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2 *>(out + r * BR_NT + 2 * ig + 16 * j) = make_double2(acc[i][2 * j], acc[i][2 * j + 1]);
    }
    out[BR_MT * BR_NT + tid] = csum;
}

// C (+)= sum over the S partial tiles, in slice order; eight CTAs per tile.
__global__ void __launch_bounds__(256)
batch_reduce_finish_kernel(const BrBatch P, int S, int accumulate, const double *__restrict__ ws) {
    const BrTile T = P.t[blockIdx.y];
    const double *base = ws + (size_t)blockIdx.y * S * BR_TILE_DOUBLES;
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

static int br_plan(const anqs_brg_problem_t *p, int n, int64_t K, std::vector<BrTile> &tiles, int &S, int64_t &k_per_split) {
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
    const int64_t want = 2 * (int64_t)sm_count_of_current_device() / (int64_t)tiles.size();   // one wave of two CTAs per SM
    const int64_t most = std::max<int64_t>(1, K / (4 * BR_KT));   // at least four K tiles per CTA
    S = (int)std::max<int64_t>(1, std::min(want, most));
    k_per_split = ((K + S - 1) / S + BR_KT - 1) / BR_KT * BR_KT;
    S = (int)((K + k_per_split - 1) / k_per_split);
    return 0;
}

}  // namespace anqs

using namespace anqs;

extern "C" int64_t anqs_batch_reduce_workspace(const anqs_brg_problem_t *problems, int n_problems, int64_t K) {
    if (!problems || n_problems <= 0 || K <= 0) return 0;
    std::vector<BrTile> tiles;
    int S;
    int64_t kps;
    if (br_plan(problems, n_problems, K, tiles, S, kps)) return -1;
    return (int64_t)tiles.size() * S * BR_TILE_DOUBLES * (int64_t)sizeof(double);
}

extern "C" int anqs_batch_reduce_gemm(const anqs_brg_problem_t *problems, int n_problems, int64_t K, int accumulate, void *workspace,
                                      int64_t workspace_bytes, void *stream) {
    ANQS_REQUIRE(problems && n_problems > 0, "no problems");
    ANQS_REQUIRE(K > 0, "empty batch");
    std::vector<BrTile> tiles;
    int S;
    int64_t kps;
    ANQS_REQUIRE(br_plan(problems, n_problems, K, tiles, S, kps) == 0, "bad problem (null pointer, N > 64 or leading dimension too small)");
    ANQS_REQUIRE(workspace && workspace_bytes >= (int64_t)tiles.size() * S * BR_TILE_DOUBLES * (int64_t)sizeof(double), "workspace too small");
    double *ws = (double *)workspace;
    for (size_t t0 = 0; t0 < tiles.size(); t0 += BR_MAX_TILES) {
        const int nt = (int)std::min<size_t>(BR_MAX_TILES, tiles.size() - t0);
        BrBatch b;
        for (int i = 0; i < nt; ++i) b.t[i] = tiles[t0 + i];
        double *w = ws + t0 * S * BR_TILE_DOUBLES;
        batch_reduce_gemm_kernel<<<dim3(S, nt), BR_THREADS, 0, (cudaStream_t)stream>>>(b, K, kps, w);
        batch_reduce_finish_kernel<<<dim3(8, nt), 256, 0, (cudaStream_t)stream>>>(b, S, accumulate, w);
    }
    ANQS_CUDA(cudaGetLastError());
    return 0;
}
