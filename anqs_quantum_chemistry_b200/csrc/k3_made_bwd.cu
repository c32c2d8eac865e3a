// Kernel family 3, backward: the per-sample chain of the MADE wave function's gradient (reference: autograd through
// ANQS:407-485 / LAP:63-163 / MLP:217-246).
//
// d log psi / d theta splits into a per-sample chain (output-layer gradient -> hidden-layer gradients, 64-wide, needs the saved
// activations and conditional probabilities of the forward pass) and reductions over the batch (outer products summed over
// samples).  made_backward_kernel does the whole chain for both sub-networks in one launch and writes exactly the operands
// of the batch reductions, which are plain GEMMs / column sums and stay library calls:
//   dY[net][B][Q*DM]   gradient w.r.t. the output layer's pre-activations
//                        log-abs network:  g_re * ([d == chosen] - p_qd)   (the mean subtraction drops out: the entries sum to 0)
//                        phase network:    pi * g_im * [d == chosen]
//   da[net][l][B][64]  gradient w.r.t. the pre-activation of hidden layer l: dh * (1 - h_l^2),
//                        dh_{l-1} = da_l W_l (+ da_l on the residual connection, MLP:237-239)
//   x[B][n]            the 1 - 2 bit input encoding (MLP:205-215)
// so that  grad W_out = dY^T h_last, grad b_out = sum_s dY, grad W_l = da_l^T (h_{l-1} | x), grad b_l = sum_s da_l.
// Same 64-sample x 64-output DFMA tiles as the forward kernel.
#include <algorithm>

#include "common.cuh"
#include "made_common.cuh"

namespace anqs {

constexpr size_t MDB_SMEM = (size_t)2 * 64 * MD_S * sizeof(double) + 64 * sizeof(uint64_t) + 64 * sizeof(double2);

// acc[ss][jj] += sum_k act[k][ty*4+ss] * wt[k][tx+16*jj]; `act` is consumed (its rows serve as the hand-over scratch of gemm_tile)
__device__ __forceinline__ void gemm_tile_acc(double *act, const double *wt, int K, int tx, int ty, double (&acc)[4][4]) {
    double p[4][4];
    gemm_tile(act, wt, K, tx, ty, p, act);
#pragma unroll
    for (int ss = 0; ss < 4; ++ss)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ss][jj] += p[ss][jj];
}

// PHASE_DY = false: the phase network's output-layer gradient (one non-zero per sample and qudit) is not written out; its
// weight gradient is the row scatter of made_phase_output_kernel below instead of a dense product with 63/64 zeros.
template <bool PHASE_DY>
__global__ void __launch_bounds__(MD_THREADS, 2)
made_backward_kernel(const anqs_made_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, const double2 *__restrict__ grad_out,
                     const double *__restrict__ save_h, const double *__restrict__ save_p, double *__restrict__ dY,
                     double *__restrict__ da_out, double *__restrict__ x_out) {
    extern __shared__ __align__(16) unsigned char md_smem[];
    double *act = reinterpret_cast<double *>(md_smem);
    double *wt = act + 64 * MD_S;
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(wt + 64 * MD_S);
    double2 *s_g = reinterpret_cast<double2 *>(s_idx + 64);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = P.qubit_num, Q = P.qudit_num, DM = P.max_qudit_dim, depth = P.depth;
    const size_t QD = (size_t)Q * DM;
    const int64_t ntiles = (B + MD_TB - 1) / MD_TB;
    const double PI = 3.14159265358979323846;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * MD_TB;
        __syncthreads();
        if (tid < 64) {
            const bool ok = base + tid < B;
            s_idx[tid] = ok ? (uint64_t)idx_in[base + tid] : 0ull;
            s_g[tid] = ok ? grad_out[base + tid] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        for (int e = tid; e < 64 * n; e += MD_THREADS) {  // input encoding, [sample][qubit]
            const int s = e / n, k = e - s * n;
            if (base + s < B) x_out[(size_t)(base + s) * n + k] = 1.0 - 2.0 * (double)((s_idx[s] >> k) & 1ull);
        }
        for (int net = 0; net < 2; ++net) {
            const double *const *Ws = net == 0 ? P.w_abs : P.w_phase;
            double dh[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dh[a][b] = 0.0;
            double *dYn = dY + (size_t)net * (size_t)B * QD;
            if (net == 0) {
                // output layer of the log-abs network: dY tile per qudit, dh += dY_q W_out[q]
                for (int q = 0; q < Q; ++q) {
                    const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
                    __syncthreads();  // the previous GEMM is done with act / wt
                    // A warp moves 4 samples x 8 outcomes at a time (full 64-byte runs of save_p / dY; the transposed stores hit
                    // (4 d + s) mod 16 = every bank pair twice), all sixteen loads of a thread in flight before the first use.
                    {
                        const int lane = tid & 31, warp = tid >> 5, dq = lane & 7, sq = lane >> 3;
                        double pv[16], wv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int t = warp + 8 * i, s = (t >> 3) * 4 + sq, d = (t & 7) * 8 + dq;
                            pv[i] = (d < DM && base + s < B) ? __ldg(save_p + ((size_t)(base + s) * Q + q) * DM + d) : 0.0;
                            const int e = tid + MD_THREADS * i, dd = e >> 6, j = e & 63;  // wt[d][j] = W[(q DM + d)][j], as the rows lie
                            wv[i] = dd < DM ? __ldg(Ws[depth] + ((size_t)q * DM + dd) * MD_W + j) : 0.0;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int t = warp + 8 * i, s = (t >> 3) * 4 + sq, d = (t & 7) * 8 + dq;
                            double v = 0.0;
                            if (d < DM && base + s < B) {
                                const int chosen = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                                v = s_g[s].x * ((d == chosen ? 1.0 : 0.0) - pv[i]);
                                dYn[(size_t)(base + s) * QD + (size_t)q * DM + d] = v;
                            }
                            act[d * MD_S + s] = v;
                            const int e = tid + MD_THREADS * i;
                            wt[(e >> 6) * MD_S + (e & 63)] = wv[i];
                        }
                    }
                    __syncthreads();
                    gemm_tile_acc(act, wt, DM, tx, ty, dh);
                }
            } else {
                // phase network: only the chosen outcome of every qudit carries a gradient (arg psi = pi * sum_q y[q, chosen])
                for (int q = 0; q < Q; ++q) {
                    const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
                    if (PHASE_DY) {
                        for (int e = tid; e < 64 * 64; e += MD_THREADS) {
                            const int s = e >> 6, d = e & 63;
                            if (d < DM && base + s < B) {
                                const int chosen = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                                dYn[(size_t)(base + s) * QD + (size_t)q * DM + d] = d == chosen ? PI * s_g[s].y : 0.0;
                            }
                        }
                    }
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        const int s = ty * 4 + ss;
                        const int chosen = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                        const double *w = Ws[depth] + ((size_t)q * DM + chosen) * MD_W;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) dh[ss][jj] += __ldg(w + tx + 16 * jj);
                    }
                }
#pragma unroll
                for (int ss = 0; ss < 4; ++ss)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) dh[ss][jj] *= PI * s_g[ty * 4 + ss].y;
            }
            // hidden layers, last to first
            for (int l = depth - 1; l >= 0; --l) {
                double da[4][4];
#pragma unroll
                for (int ss = 0; ss < 4; ++ss) {
                    const int64_t row = base + ty * 4 + ss;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = tx + 16 * jj;
                        double v = 0.0;
                        if (row < B) {
                            const double h = __ldg(save_h + (((size_t)net * depth + l) * (size_t)B + (size_t)row) * MD_W + j);
                            v = dh[ss][jj] * (1.0 - h * h);
                            da_out[(((size_t)net * depth + l) * (size_t)B + (size_t)row) * MD_W + j] = v;
                        }
                        da[ss][jj] = v;
                    }
                }
                if (l > 0) {
                    __syncthreads();
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) act[(tx + 16 * jj) * MD_S + ty * 4 + ss] = da[ss][jj];
                    for (int e = tid; e < 64 * 64; e += MD_THREADS) {
                        const int j = e >> 6, i = e & 63;
                        wt[j * MD_S + i] = __ldg(Ws[l] + (size_t)j * MD_W + i);
                    }
                    __syncthreads();
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dh[a][b] = P.use_res ? da[a][b] : 0.0;  // residual connection (MLP:237-239)
                    gemm_tile_acc(act, wt, MD_W, tx, ty, dh);
                }
            }
        }
    }
}

// Output-layer gradient of the phase network (arg psi = pi * sum_q y[q, chosen_q], LAP:97, ANQS:450-454):
//   grad W_out[q DM + d][j] = sum_s [chosen_q(s) = d] pi g_s h_s[j],   grad b_out[q DM + d] = sum_s [chosen_q(s) = d] pi g_s
// - one row of 64 numbers per (sample, qudit), where the dense product dY^T h multiplies 63 zeros for every one of them.
// A CTA owns a slice of the samples; for one qudit at a time each of its warps adds its samples' rows into a private
// [DM][64] (+ [DM] bias) tile in shared memory, walking the samples in order; the warps' tiles are then added in warp order
// into the CTA's partial result, and made_phase_output_finish_kernel adds the partial results in CTA order: deterministic.
// The hidden rows are re-read once per qudit (from L2 when the chunk fits).
constexpr int PO_WARPS = 6, PO_TILE = 64 * 64 + 64;   // 6 x 33 KB of accumulator tiles per SM
constexpr size_t PO_SMEM = (size_t)PO_WARPS * PO_TILE * sizeof(double);

__global__ void __launch_bounds__(PO_WARPS * 32, 1)
made_phase_output_kernel(const anqs_made_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, const double2 *__restrict__ grad_out,
                         const double *__restrict__ h_last, double *__restrict__ partial) {
    extern __shared__ __align__(16) unsigned char po_smem[];
    double *tiles = reinterpret_cast<double *>(po_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *acc = tiles + (size_t)warp * PO_TILE;
    const int Q = P.qudit_num, DM = P.max_qudit_dim;
    const double PI = 3.14159265358979323846;
    const int64_t per = (B + gridDim.x - 1) / gridDim.x;
    const int64_t s0 = (int64_t)blockIdx.x * per, s1 = min(B, s0 + per);
    for (int q = 0; q < Q; ++q) {
        const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
        const uint64_t omask = (1ull << bits) - 1ull;
        for (int e = lane; e < PO_TILE; e += 32) acc[e] = 0.0;
        __syncwarp();
        int64_t s = s0 + warp;
        for (; s + 3 * PO_WARPS < s1; s += 4 * PO_WARPS) {   // four samples per step: twelve loads in flight per lane
            uint64_t x[4];
            double g[4], h0[4], h1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t r = s + u * PO_WARPS;
                x[u] = (uint64_t)__ldg(idx_in + r);
                g[u] = PI * __ldg(&grad_out[r].y);
                h0[u] = __ldg(h_last + r * MD_W + lane);
                h1[u] = __ldg(h_last + r * MD_W + 32 + lane);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = (int)((x[u] >> start) & omask);
                acc[c * 64 + lane] = fma(g[u], h0[u], acc[c * 64 + lane]);
                acc[c * 64 + 32 + lane] = fma(g[u], h1[u], acc[c * 64 + 32 + lane]);
                if (lane == 0) acc[64 * 64 + c] += g[u];
            }
        }
        for (; s < s1; s += PO_WARPS) {
            const uint64_t x = (uint64_t)__ldg(idx_in + s);
            const double g = PI * __ldg(&grad_out[s].y);
            const int c = (int)((x >> start) & omask);
            acc[c * 64 + lane] = fma(g, __ldg(h_last + s * MD_W + lane), acc[c * 64 + lane]);
            acc[c * 64 + 32 + lane] = fma(g, __ldg(h_last + s * MD_W + 32 + lane), acc[c * 64 + 32 + lane]);
            if (lane == 0) acc[64 * 64 + c] += g;
        }
        __syncthreads();
        double *out = partial + ((size_t)blockIdx.x * Q + q) * PO_TILE;
        for (int e = tid; e < PO_TILE; e += PO_WARPS * 32) {
            double v = tiles[e];
#pragma unroll
            for (int w = 1; w < PO_WARPS; ++w) v += tiles[(size_t)w * PO_TILE + e];
            out[e] = v;
        }
        __syncthreads();
        (void)DM;
    }
}

__global__ void made_phase_output_finish_kernel(const double *__restrict__ partial, int n_cta, int Q, int DM, int accumulate,
                                                double *__restrict__ gW, double *__restrict__ gb) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // over Q * PO_TILE
    if (e >= Q * PO_TILE) return;
    const int q = e / PO_TILE, r = e - q * PO_TILE;
    const bool is_bias = r >= 64 * 64;
    const int d = is_bias ? r - 64 * 64 : r >> 6, j = r & 63;
    if (d >= DM || (is_bias && gb == nullptr)) return;
    double v = 0.0;
    for (int c = 0; c < n_cta; ++c) v += partial[((size_t)c * Q + q) * PO_TILE + r];
    double *dst = is_bias ? gb + (size_t)q * DM + d : gW + ((size_t)q * DM + d) * MD_W + j;
    *dst = accumulate ? *dst + v : v;
}

// NADE mode: the same chain for every (sub-network, qudit) MLP (LAP:24-42): output width D_q = 2^bits of the qudit, first layer
// fed by the qubits before the qudit.  Outputs: dY[2][B][Q*DM] (entries d >= D_q are zero), da[2][Q][depth][B][64], x[B][n].
__global__ void __launch_bounds__(MD_THREADS, 2)
nade_backward_kernel(const anqs_nade_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, const double2 *__restrict__ grad_out,
                     const double *__restrict__ save_h, const double *__restrict__ save_p, double *__restrict__ dY,
                     double *__restrict__ da_out, double *__restrict__ x_out) {
    extern __shared__ __align__(16) unsigned char md_smem[];
    double *act = reinterpret_cast<double *>(md_smem);
    double *wt = act + 64 * MD_S;
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(wt + 64 * MD_S);
    double2 *s_g = reinterpret_cast<double2 *>(s_idx + 64);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = P.qubit_num, Q = P.qudit_num, DM = P.max_qudit_dim, depth = P.depth;
    const size_t QD = (size_t)Q * DM;
    const int64_t ntiles = (B + MD_TB - 1) / MD_TB;
    const double PI = 3.14159265358979323846;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * MD_TB;
        __syncthreads();
        if (tid < 64) {
            const bool ok = base + tid < B;
            s_idx[tid] = ok ? (uint64_t)idx_in[base + tid] : 0ull;
            s_g[tid] = ok ? grad_out[base + tid] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        for (int e = tid; e < 64 * n; e += MD_THREADS) {
            const int s = e / n, k = e - s * n;
            if (base + s < B) x_out[(size_t)(base + s) * n + k] = 1.0 - 2.0 * (double)((s_idx[s] >> k) & 1ull);
        }
        for (int net = 0; net < 2; ++net) {
            double *dYn = dY + (size_t)net * (size_t)B * QD;
            for (int q = 0; q < Q; ++q) {
                const double *const *tab = P.ptrs + ((size_t)(net * Q + q) * (depth + 1)) * 2;
                const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start, D = 1 << bits;
                const double *W_out = tab[2 * depth];
                double dh[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dh[a][b] = 0.0;
                __syncthreads();  // the previous GEMM is done with act / wt
                {   // same staging as made_backward_kernel: 4 samples x 8 outcomes per warp step, loads first
                    const int lane = tid & 31, warp = tid >> 5, dq = lane & 7, sq = lane >> 3;
                    double pv[16], wv[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int t = warp + 8 * i, s = (t >> 3) * 4 + sq, d = (t & 7) * 8 + dq;
                        pv[i] = (net == 0 && d < D && base + s < B) ? __ldg(save_p + ((size_t)(base + s) * Q + q) * DM + d) : 0.0;
                        const int e = tid + MD_THREADS * i, dd = e >> 6, j = e & 63;
                        wv[i] = dd < D ? __ldg(W_out + (size_t)dd * MD_W + j) : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int t = warp + 8 * i, s = (t >> 3) * 4 + sq, d = (t & 7) * 8 + dq;
                        double v = 0.0;
                        if (d < DM && base + s < B) {
                            const int chosen = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                            if (d < D)
                                v = net == 0 ? s_g[s].x * ((d == chosen ? 1.0 : 0.0) - pv[i]) : (d == chosen ? PI * s_g[s].y : 0.0);
                            dYn[(size_t)(base + s) * QD + (size_t)q * DM + d] = v;
                        }
                        act[d * MD_S + s] = v;
                        const int e = tid + MD_THREADS * i;
                        wt[(e >> 6) * MD_S + (e & 63)] = wv[i];
                    }
                }
                __syncthreads();
                gemm_tile_acc(act, wt, D, tx, ty, dh);
                for (int l = depth - 1; l >= 0; --l) {
                    double da[4][4];
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        const int64_t row = base + ty * 4 + ss;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = tx + 16 * jj;
                            double v = 0.0;
                            if (row < B) {
                                const size_t off = ((((size_t)net * Q + q) * depth + l) * (size_t)B + (size_t)row) * MD_W + j;
                                const double h = __ldg(save_h + off);
                                v = dh[ss][jj] * (1.0 - h * h);
                                da_out[off] = v;
                            }
                            da[ss][jj] = v;
                        }
                    }
                    if (l > 0) {
                        __syncthreads();
#pragma unroll
                        for (int ss = 0; ss < 4; ++ss)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) act[(tx + 16 * jj) * MD_S + ty * 4 + ss] = da[ss][jj];
                        const double *W = tab[2 * l];
                        for (int e = tid; e < 64 * 64; e += MD_THREADS) {
                            const int j = e >> 6, i = e & 63;
                            wt[j * MD_S + i] = __ldg(W + (size_t)j * MD_W + i);
                        }
                        __syncthreads();
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) dh[a][b] = P.use_res ? da[a][b] : 0.0;
                        gemm_tile_acc(act, wt, MD_W, tx, ty, dh);
                    }
                }
            }
        }
    }
}

}  // namespace anqs

using namespace anqs;

static int made_chain_launch(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out, const double *d_save_h,
                             const double *d_save_p, double *d_dY, double *d_da, double *d_x, bool phase_dy, void *stream) {
    ANQS_REQUIRE(desc, "null network descriptor");
    ANQS_REQUIRE(desc->width == MD_W && desc->max_qudit_dim <= 64 && desc->depth >= 1 && desc->depth <= 4, "unsupported network shape");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_grad_out && d_save_h && d_save_p && d_dY && d_da && d_x, "null pointer");
    auto kern = phase_dy ? made_backward_kernel<true> : made_backward_kernel<false>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MDB_SMEM));
    const int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, MD_THREADS, MDB_SMEM, (cudaStream_t)stream>>>(*desc, d_idx, n, (const double2 *)d_grad_out, d_save_h, d_save_p, d_dY, d_da, d_x);
    ANQS_LAUNCH_CHECK();
    return 0;
}

extern "C" int anqs_made_backward_chain(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                        const double *d_save_h, const double *d_save_p, double *d_dY, double *d_da, double *d_x,
                                        void *stream) {
    return made_chain_launch(desc, d_idx, n, d_grad_out, d_save_h, d_save_p, d_dY, d_da, d_x, true, stream);
}

extern "C" int anqs_made_backward_chain_abs(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                            const double *d_save_h, const double *d_save_p, double *d_dY_abs, double *d_da, double *d_x,
                                            void *stream) {
    return made_chain_launch(desc, d_idx, n, d_grad_out, d_save_h, d_save_p, d_dY_abs, d_da, d_x, false, stream);
}

extern "C" int64_t anqs_made_phase_output_workspace(const anqs_made_desc_t *desc) {
    if (!desc || desc->qudit_num < 1) return -1;
    return (int64_t)sm_count_of_current_device() * desc->qudit_num * PO_TILE * (int64_t)sizeof(double);
}

extern "C" int anqs_made_phase_output_grad(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                           const double *d_h_last, int accumulate, double *d_gW, double *d_gb, void *d_work,
                                           int64_t work_bytes, void *stream) {
    ANQS_REQUIRE(desc, "null network descriptor");
    ANQS_REQUIRE(desc->width == MD_W && desc->max_qudit_dim <= 64, "unsupported network shape");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_grad_out && d_h_last && d_gW && d_work, "null pointer");
    ANQS_REQUIRE(work_bytes >= anqs_made_phase_output_workspace(desc), "workspace too small");
    const int grid = (int)std::min<int64_t>((n + PO_WARPS - 1) / PO_WARPS, (int64_t)sm_count_of_current_device());
    cudaStream_t s = (cudaStream_t)stream;
    ANQS_CUDA(cudaFuncSetAttribute(made_phase_output_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PO_SMEM));
    made_phase_output_kernel<<<grid, PO_WARPS * 32, PO_SMEM, s>>>(*desc, d_idx, n, (const double2 *)d_grad_out, d_h_last, (double *)d_work);
    ANQS_LAUNCH_CHECK();
    const int total = desc->qudit_num * PO_TILE;
    made_phase_output_finish_kernel<<<(total + 255) / 256, 256, 0, s>>>((const double *)d_work, grid, desc->qudit_num, desc->max_qudit_dim,
                                                                        accumulate, d_gW, d_gb);
    ANQS_LAUNCH_CHECK();
    return 0;
}

extern "C" int anqs_nade_backward_chain(const anqs_nade_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                        const double *d_save_h, const double *d_save_p, double *d_dY, double *d_da, double *d_x,
                                        void *stream) {
    ANQS_REQUIRE(desc, "null network descriptor");
    ANQS_REQUIRE(desc->width == MD_W && desc->max_qudit_dim <= 64 && desc->depth >= 1 && desc->depth <= 4 && desc->ptrs, "unsupported network shape");
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_grad_out && d_save_h && d_save_p && d_dY && d_da && d_x, "null pointer");
    ANQS_CUDA(cudaFuncSetAttribute(nade_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MDB_SMEM));
    const int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    nade_backward_kernel<<<grid, MD_THREADS, MDB_SMEM, (cudaStream_t)stream>>>(*desc, d_idx, n, (const double2 *)d_grad_out, d_save_h,
                                                                              d_save_p, d_dY, d_da, d_x);
    ANQS_LAUNCH_CHECK();
    return 0;
}
