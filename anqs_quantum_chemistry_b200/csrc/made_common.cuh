// fp64 building blocks shared by the MADE (k3_made.cu) and transformer (k5_transformer.cu) kernels: 64 x 64 x K DFMA
// tiles with operands in shared memory ([k][row] layout, row stride MD_S), 16-lane row reductions, and the memo index of
// the LocallyDecomposableMasker (reference MSK:67-73 on MSK:156-167).
#pragma once
#include "common.cuh"

namespace anqs {

constexpr int MD_TB = 64;        // samples per tile
constexpr int MD_W = 64;         // hidden width (reference default, MLP:17-23) and max outcomes per qudit
constexpr int MD_S = 68;         // shared-memory row stride in doubles (16-byte aligned rows, conflict-light stores)
constexpr int MD_THREADS = 256;  // 16 x 16 threads, 4 x 4 outputs each

__device__ __forceinline__ long long floor_div(long long a, long long b) {
    long long q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// memo index of the quantum numbers accumulated over the bits of `prefix` (MSK:67-73 on MSK:156-167)
__device__ __forceinline__ long long memo_index(const anqs_made_desc_t &P, uint64_t prefix) {
    long long idx = 0;
    for (int s = 0; s < P.sym_num; ++s) {
        const int64_t *d = P.sym[s];
        long long e;
        if (d[0] == 0)
            e = d[7] + __popcll(prefix & (uint64_t)d[1]) - __popcll(prefix & (uint64_t)d[2]);
        else
            e = (__popcll(prefix & (uint64_t)d[1]) & 1) ? -d[7] : d[7];
        idx += floor_div(e * d[3] + d[4], d[5]) * d[6];
    }
    return idx;
}

// wt[k][j] = W[(row0 + j) * K + k] for j < rows, 0 otherwise  (nn.Linear layout [out][in])
// A warp moves 8 rows x 4 consecutive k at a time: every row contributes one full 32-byte sector of W, and the 32 stores
// land on (4 k + j) mod 16 = every 8-byte bank pair exactly twice (MD_S = 68 = 4 mod 16), the two-wavefront minimum.
__device__ __forceinline__ void load_weights_t(double *wt, const double *__restrict__ W, int row0, int rows, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, jo = lane >> 2;
    // K <= 64: at most 16 steps per warp; all loads of a thread are issued before the first store
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        v[i] = (k < K && j < rows) ? __ldg(W + (size_t)(row0 + j) * K + k) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        if (k < K) wt[k * MD_S + j] = v[i];
    }
}

// acc[ss][jj] = sum_k act[k][ty*4+ss] * wt[k][tx+16*jj]
__device__ __forceinline__ void gemm_tile(const double *act, const double *wt, int K, int tx, int ty, double (&acc)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const double2 a01 = *reinterpret_cast<const double2 *>(act + k * MD_S + ty * 4);
        const double2 a23 = *reinterpret_cast<const double2 *>(act + k * MD_S + ty * 4 + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        double w[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) w[jj] = wt[k * MD_S + tx + 16 * jj];
#pragma unroll
        for (int ss = 0; ss < 4; ++ss)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[ss][jj] = fma(a[ss], w[jj], acc[ss][jj]);
    }
}

__device__ __forceinline__ double row_sum16(double v) {
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double row_max16(double v) {
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}


// same as memo_index for descriptors other than anqs_made_desc_t
__device__ __forceinline__ long long memo_index_of(int sym_num, const int64_t (*sym)[8], uint64_t prefix) {
    long long idx = 0;
    for (int s = 0; s < sym_num; ++s) {
        const int64_t *d = sym[s];
        long long e;
        if (d[0] == 0)
            e = d[7] + __popcll(prefix & (uint64_t)d[1]) - __popcll(prefix & (uint64_t)d[2]);
        else
            e = (__popcll(prefix & (uint64_t)d[1]) & 1) ? -d[7] : d[7];
        idx += floor_div(e * d[3] + d[4], d[5]) * d[6];
    }
    return idx;
}

}  // namespace anqs
