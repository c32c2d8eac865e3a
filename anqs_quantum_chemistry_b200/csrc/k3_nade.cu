// Kernel family 3, NADE mode: one pair of MLPs per qudit instead of one masked network
// (reference ANQS:410-428 with LAP:24-42, 63-103, 114-134 and MLP:217-246 for is_made=False).
//
// Qudit q has its own log-abs and phase MLP: input = the q*k bits before it as 1 - 2 bit (for the first qudit the
// reference feeds the constant 0.5 through the same encoding, i.e. the input 0: MLP:205-215), `depth` tanh hidden layers
// of width 64 with residual adds on layers 1..depth-1, linear output of width qudit_dims[q].  The conditional
// log-amplitudes get their mean over the qudit's own outcomes subtracted (LAP:118-119 - not over max_qudit_dim as in MADE
// mode), are masked by the symmetry continuation mask and normalised (ANQS:392-405); the phase head is evaluated only at the
// chosen outcome.  Same tiling as made_forward_kernel: 64 samples per tile, every layer one 64 x 64 x K DFMA tile.
//
// The per-qudit weights come through a device table of pointers: entry ((net * Q + q) * (depth + 1) + layer) * 2 + {0: weight,
// 1: bias}, net 0 = log_abs_subnet[q], net 1 = phase_subnet[q].
#include <algorithm>

#include "common.cuh"
#include "made_common.cuh"

namespace anqs {

constexpr size_t ND_SMEM = (size_t)3 * 64 * MD_S * sizeof(double) + 2 * 64 * sizeof(uint64_t) + 4 * 64 * sizeof(double);

template <int MODE>  // 0: log psi, 1: conditional log|psi| of qudit level_q
__global__ void __launch_bounds__(MD_THREADS, 2)
nade_forward_kernel(const anqs_nade_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, int level_q,
                    double2 *__restrict__ log_psi, double *__restrict__ cond_out, double *__restrict__ save_h,
                    double *__restrict__ save_p) {
    extern __shared__ __align__(16) unsigned char nd_smem[];
    double *act0 = reinterpret_cast<double *>(nd_smem);
    double *act1 = act0 + 64 * MD_S;
    double *wt = act1 + 64 * MD_S;
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(wt + 64 * MD_S);
    uint64_t *s_mask = s_idx + 64;
    double *s_im = reinterpret_cast<double *>(s_mask + 64);  // [4][64] phase partial sums

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int Q = P.qudit_num, DM = P.max_qudit_dim, depth = P.depth;
    const int64_t ntiles = (B + MD_TB - 1) / MD_TB;
    const int q_lo = MODE == 1 ? level_q : 0, q_hi = MODE == 1 ? level_q + 1 : Q;
    const int nets = MODE == 1 ? 1 : 2;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * MD_TB;
        __syncthreads();
        if (tid < 64) s_idx[tid] = base + tid < B ? (uint64_t)idx_in[base + tid] : 0ull;
        double out_re[4] = {0.0, 0.0, 0.0, 0.0};
        double im_part = 0.0;  // threads 0..63: phase of sample tid (log-psi mode)
        for (int net = 0; net < nets; ++net) {
            for (int q = q_lo; q < q_hi; ++q) {
                const double *const *tab = P.ptrs + ((size_t)(net * Q + q) * (depth + 1)) * 2;
                const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start, D = 1 << bits;
                const int K0 = start == 0 ? 1 : start;  // LAP:26: in_num = 1 for the first qudit
                __syncthreads();
                for (int e = tid; e < K0 * 64; e += MD_THREADS) {
                    const int k = e >> 6, s = e & 63;
                    act0[k * MD_S + s] = start == 0 ? 0.0 : 1.0 - 2.0 * (double)((s_idx[s] >> k) & 1ull);
                }
                double *cur = act0, *nxt = act1;
                for (int l = 0; l < depth; ++l) {
                    const int K = l == 0 ? K0 : MD_W;
                    const double *W = tab[2 * l], *bvec = tab[2 * l + 1];
                    load_weights_t(wt, W, 0, MD_W, K);
                    __syncthreads();
                    double acc[4][4];
                    gemm_tile(cur, wt, K, tx, ty, acc, nxt);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = tx + 16 * jj;
                        const double bias = bvec ? __ldg(bvec + j) : 0.0;
#pragma unroll
                        for (int ss = 0; ss < 4; ++ss) {
                            double v = acc[ss][jj] + bias;
                            if (P.use_res && l > 0) v += cur[j * MD_S + ty * 4 + ss];  // MLP:237-239
                            v = tanh(v);
                            nxt[j * MD_S + ty * 4 + ss] = v;
                            if (save_h && base + ty * 4 + ss < B)
                                save_h[((((size_t)net * Q + q) * depth + l) * (size_t)B + (size_t)(base + ty * 4 + ss)) * MD_W + j] = v;
                        }
                    }
                    __syncthreads();
                    double *t = cur;
                    cur = nxt;
                    nxt = t;
                }
                const double *W3 = tab[2 * depth], *b3 = tab[2 * depth + 1];
                if (net == 0) {
                    load_weights_t(wt, W3, 0, D, MD_W);
                    if (tid < 64) {
                        const uint64_t x = s_idx[tid];
                        const uint64_t prefix = start == 0 ? 0ull : (x & ((1ull << start) - 1ull));
                        uint64_t mw;
                        if (P.du[q]) {
                            mw = D >= 64 ? ~0ull : ((1ull << D) - 1ull);
                        } else {
                            const long long mi = memo_index_of(P.sym_num, P.sym, prefix);
                            mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)q * P.memo_size + mi) : 0ull;
                        }
                        s_mask[tid] = mw;
                    }
                    __syncthreads();
                    double acc[4][4];
                    gemm_tile(cur, wt, MD_W, tx, ty, acc, nxt);
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        const int s = ty * 4 + ss;
                        const uint64_t mw = s_mask[s];
                        const int chosen = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                        double z[4];
                        double sum = 0.0;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int d = tx + 16 * jj;
                            z[jj] = acc[ss][jj] + ((b3 && d < D) ? __ldg(b3 + d) : 0.0);
                            sum += d < D ? z[jj] : 0.0;
                        }
                        if (P.subtract_mean) {  // over the qudit's own D outcomes, before masking (LAP:118-119)
                            const double mean = row_sum16(sum) / (double)D;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) z[jj] -= mean;
                        }
                        double mx = -INFINITY;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int d = tx + 16 * jj;
                            if (d < D && ((mw >> d) & 1ull)) mx = fmax(mx, z[jj]);
                        }
                        mx = row_max16(mx);
                        double se = 0.0, ez[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int d = tx + 16 * jj;
                            ez[jj] = (d < D && ((mw >> d) & 1ull)) ? exp(2.0 * (z[jj] - mx)) : 0.0;
                            se += ez[jj];
                        }
                        se = row_sum16(se);
                        const double L = mx + 0.5 * log(se);
                        const bool any = mw != 0ull && mx > -INFINITY;
                        if (MODE == 1) {
                            if (base + s < B) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int d = tx + 16 * jj;
                                    if (d < DM)
                                        cond_out[(size_t)(base + s) * DM + d] =
                                            (any && d < D && ((mw >> d) & 1ull)) ? z[jj] - L : -INFINITY;
                                }
                            }
                        } else {
                            double pick = 0.0;
                            const double inv_se = 1.0 / se;   // p_d = exp(2 (z_d - max)) / sum: no second exponential
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int d = tx + 16 * jj;
                                const bool allowed = any && d < D && ((mw >> d) & 1ull);
                                if (d == chosen) pick = allowed ? z[jj] - L : -INFINITY;
                                if (save_p && d < DM && base + s < B)
                                    save_p[((size_t)(base + s) * Q + q) * DM + d] = allowed ? ez[jj] * inv_se : 0.0;
                            }
                            out_re[ss] += row_sum16(pick);
                        }
                    }
                } else {
                    // phase network: only the row of the chosen outcome (LAP:97, ANQS:419-424)
                    const int s = tid & 63, g = tid >> 6;
                    const int row = (int)((s_idx[s] >> start) & ((1ull << bits) - 1ull));
                    const double *w = W3 + (size_t)row * MD_W;
                    double dot = (g == 0 && b3) ? __ldg(b3 + row) : 0.0;
                    for (int k = g * 16; k < g * 16 + 16; ++k) dot = fma(__ldg(w + k), cur[k * MD_S + s], dot);
                    __syncthreads();
                    s_im[g * 64 + s] = dot;
                    __syncthreads();
                    if (tid < 64) im_part += s_im[tid] + s_im[64 + tid] + s_im[128 + tid] + s_im[192 + tid];
                }
            }
        }
        if (MODE == 0) {
            double *s_re = wt;
            __syncthreads();
            if (tx == 0) {
#pragma unroll
                for (int ss = 0; ss < 4; ++ss) s_re[ty * 4 + ss] = out_re[ss];
            }
            __syncthreads();
            if (tid < 64 && base + tid < B) {
                const double re = s_re[tid];
                log_psi[base + tid] = make_double2(re, re == -INFINITY ? 0.0 : 3.14159265358979323846 * im_part);
            }
        }
    }
}

}  // namespace anqs

using namespace anqs;

static int nade_check(const anqs_nade_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->qudit_num >= 1 && P->qudit_num <= 64, "qudit_num must be in [1, 64]");
    ANQS_REQUIRE(P->max_qudit_dim >= 2 && P->max_qudit_dim <= 64, "max_qudit_dim must be in [2, 64]");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] hidden layers");
    ANQS_REQUIRE(P->width == MD_W, "hidden width must be 64 (the reference default)");
    ANQS_REQUIRE(P->sym_num >= 0 && P->sym_num <= 8, "at most 8 symmetries");
    ANQS_REQUIRE(P->qudit_starts[0] == 0 && P->qudit_starts[P->qudit_num] == P->qubit_num, "qudit_starts must span the qubits");
    ANQS_REQUIRE(P->ptrs, "null weight-pointer table");
    ANQS_REQUIRE(P->cont_mask && P->memo_size >= 1, "null continuation-mask table");
    return 0;
}

extern "C" {

int anqs_nade_log_psi(const anqs_nade_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, double *d_save_h,
                      double *d_save_p, void *stream) {
    if (nade_check(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_log_psi, "null pointer");
    auto kern = nade_forward_kernel<0>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ND_SMEM));
    const int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, MD_THREADS, ND_SMEM, (cudaStream_t)stream>>>(*desc, d_idx, n, 0, (double2 *)d_log_psi, nullptr, d_save_h, d_save_p);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_nade_cond_log_abs(const anqs_nade_desc_t *desc, int qudit_idx, const int64_t *d_prefix, int64_t n, double *d_cond,
                           void *stream) {
    if (nade_check(desc)) return 1;
    ANQS_REQUIRE(qudit_idx >= 0 && qudit_idx < desc->qudit_num, "qudit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_prefix && d_cond, "null pointer");
    auto kern = nade_forward_kernel<1>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ND_SMEM));
    const int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, MD_THREADS, ND_SMEM, (cudaStream_t)stream>>>(*desc, d_prefix, n, qudit_idx, nullptr, d_cond, nullptr, nullptr);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
