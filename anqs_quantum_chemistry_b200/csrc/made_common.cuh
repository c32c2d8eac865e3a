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
__device__ __forceinline__ void fetch_weights_t(double (&v)[16], const double *__restrict__ W, int row0, int rows, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, jo = lane >> 2;
    // K <= 64: at most 16 steps per warp; all loads of a thread are issued before the first store
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        v[i] = (k < K && j < rows) ? __ldg(W + (size_t)(row0 + j) * K + k) : 0.0;
    }
}
__device__ __forceinline__ void commit_weights_t(double *wt, const double (&v)[16], int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, jo = lane >> 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        if (k < K) wt[k * MD_S + j] = v[i];
    }
}
// The two halves can be pulled apart by a caller that has other work between them: fetch the NEXT tile's weights before the
// current product, commit them after it (made_forward_kernel, transformer_backward_kernel).
__device__ __forceinline__ void load_weights_t(double *wt, const double *__restrict__ W, int row0, int rows, int K) {
    double v[16];
    fetch_weights_t(v, W, row0, rows, K);
    commit_weights_t(wt, v, K);
}

// acc[ss][jj] = sum_k act[k][ty*4+ss] * wt[k][tx+16*jj]   (a 64 x 64 x K tile per CTA of 256 threads)
//
// The products run on the FP64 tensor-core path: mma.sync m8n8k4 f64 reaches the same 37 TFLOP/s as DFMA on B200
// (scripts/microbench_dmma.cu) with one warp instruction per 256 multiply-adds instead of one per 32, and its operands come
// from 9 conflict-free 8-byte shared-memory loads per 4 k (row stride MD_S = 68 = 4 mod 16: the half-warp's addresses
// 4 * (lane % 4) + lane / 4 cover the 16 bank pairs) instead of 24 loads.  A warp multiplies its own 8 rows (samples
// 8 warp .. 8 warp + 7, the rows ty * 4 + ss of its threads) by all 64 columns: 8 accumulator fragments.  The fragments
// (lane: row lane / 4, columns 8 t + 2 (lane % 4) + {0, 1}) are then handed to the layout every epilogue is written in
// (thread: rows ty * 4 + ss, columns tx + 16 jj) through the warp's own 8 rows of `scratch`, a [64][MD_S] buffer the caller
// can spare at this point - the destination of the epilogue, or `act` itself when its values are not needed afterwards (only
// this warp reads rows 8 warp .. 8 warp + 7 of `act`).  No block-wide barrier is involved.
__device__ __forceinline__ void mma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void gemm_tile(const double *act, const double *wt, int K, int tx, int ty, double (&acc)[4][4], double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fk = lane & 3;   // A fragment: (row fr, k fk); B fragment: (k fk, column fr)
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
    const double *ap = act + fk * MD_S + warp * 8 + fr;
    const double *bp = wt + fk * MD_S + fr;
    const int Kfull = K & ~3;
#pragma unroll 2
    for (int k0 = 0; k0 < Kfull; k0 += 4) {
        const double a = ap[k0 * MD_S];
        double b[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) b[t] = bp[k0 * MD_S + 8 * t];
#pragma unroll
        for (int t = 0; t < 8; ++t) mma_m8n8k4(c[t], a, b[t]);
    }
    if (Kfull < K) {   // ragged K (first MADE layer of an odd qubit count, narrow NADE outputs): zero operands beyond K
        const bool ok = Kfull + fk < K;
        const double a = ok ? ap[Kfull * MD_S] : 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) mma_m8n8k4(c[t], a, ok ? bp[Kfull * MD_S + 8 * t] : 0.0);
    }
    __syncwarp();   // scratch may be act: every lane of the warp has read its operands
    double *sp = scratch + (2 * fk) * MD_S + warp * 8 + fr;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        sp[(8 * t) * MD_S] = c[t][0];
        sp[(8 * t + 1) * MD_S] = c[t][1];
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const double2 v01 = *reinterpret_cast<const double2 *>(scratch + (tx + 16 * jj) * MD_S + ty * 4);
        const double2 v23 = *reinterpret_cast<const double2 *>(scratch + (tx + 16 * jj) * MD_S + ty * 4 + 2);
        acc[0][jj] = v01.x;
        acc[1][jj] = v01.y;
        acc[2][jj] = v23.x;
        acc[3][jj] = v23.y;
    }
    __syncwarp();   // the epilogue may write these positions
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
