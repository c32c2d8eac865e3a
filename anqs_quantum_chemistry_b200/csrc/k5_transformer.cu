// Kernel family 5: autoregressive transformer wave function (BASELINE config 3: per-qubit conditionals from a causal
// transformer, particle-number / S_z masks).  Architecture = the reference's TransformerMADE
// (stochastic/ansatzes/legacy/anqs_primitives/made/transformer_made.py:9-48): token embedding (3 x dim, token 2 = BOS) +
// positional embedding, `depth` post-norm nn.TransformerEncoderLayer blocks (multi-head causal self-attention,
// feed-forward dim -> dim -> dim with ReLU, LayerNorm after each residual add, eps 1e-5), linear decoder to 4 numbers per
// position = (re, im) of the unnormalised conditional log-amplitude of outcome 0 and of outcome 1
// (legacy/made/real_log_psi_transformer_made.py:42-58).  The masked normalisation re -= 0.5 logsumexp(2 re) over the
// outcomes the symmetries allow (ANQS:392-405 with QG:199-213 at one qubit per qudit) and the gather of the chosen
// outcome are fused behind the decoder.
//
// fp64 parity mode (the reference is float64): dim = 64; a 64-row tile holds floor(64 / T) samples of T tokens each; every
// projection is one 64 x 64 x 64 DFMA tile (made_common.cuh) with the nn.Linear weights staged transposed in shared
// memory; attention is one thread per (token, head): two passes over the <= T causal keys (max, then exp-weighted sum),
// everything in registers; LayerNorm row statistics are 16-lane shuffles in the GEMM's own thread layout.
#include <algorithm>

#include "common.cuh"
#include "made_common.cuh"

namespace anqs {

constexpr int TF_D = 64;
constexpr size_t TF_SMEM = (size_t)5 * 64 * MD_S * sizeof(double) + 64 * 4 * sizeof(double) + 64 * sizeof(uint64_t);

// acc (+ bias) for the thread's 4 x 4 outputs, stored as out[col][row]
__device__ __forceinline__ void store_tile(double *out, const double (&acc)[4][4], const double *__restrict__ bias, int tx, int ty) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = tx + 16 * jj;
        const double b = bias ? __ldg(bias + j) : 0.0;
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) out[j * MD_S + ty * 4 + ss] = acc[ss][jj] + b;
    }
}

// x <- LayerNorm(x + acc + bias) row-wise over the 64 columns (biased variance, eps inside the square root)
__device__ __forceinline__ void residual_layer_norm(double *x, const double (&acc)[4][4], const double *__restrict__ bias,
                                                    const double *__restrict__ gamma, const double *__restrict__ beta, double eps,
                                                    int tx, int ty) {
    double v[4][4];
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) {
        double sum = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tx + 16 * jj;
            v[ss][jj] = acc[ss][jj] + (bias ? __ldg(bias + j) : 0.0) + x[j * MD_S + ty * 4 + ss];
            sum += v[ss][jj];
        }
        const double mean = row_sum16(sum) * (1.0 / 64.0);
        double sq = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            v[ss][jj] -= mean;
            sq += v[ss][jj] * v[ss][jj];
        }
        const double rstd = 1.0 / sqrt(row_sum16(sq) * (1.0 / 64.0) + eps);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tx + 16 * jj;
            x[j * MD_S + ty * 4 + ss] = v[ss][jj] * rstd * __ldg(gamma + j) + __ldg(beta + j);
        }
    }
}

// MODE 0: log psi of whole configurations.  MODE 1: normalised conditional log|psi| of qubit `level` for prefixes.
template <int MODE>
__global__ void __launch_bounds__(MD_THREADS, 1)
transformer_forward_kernel(const anqs_transformer_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, int level,
                           double2 *__restrict__ log_psi, double *__restrict__ cond_out) {
    extern __shared__ __align__(16) unsigned char tf_smem[];
    double *X = reinterpret_cast<double *>(tf_smem);   // residual stream  [64 cols][rows]
    double *Qb = X + 64 * MD_S;                        // queries, then attention output
    double *Kb = Qb + 64 * MD_S;                       // keys, then feed-forward hidden
    double *Vb = Kb + 64 * MD_S;                       // values
    double *wt = Vb + 64 * MD_S;                       // transposed weight tile
    double *s_dec = wt + 64 * MD_S;                    // decoder output [rows][4]
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(s_dec + 64 * 4);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = P.qubit_num, H = P.head_num, hd = TF_D / H;
    const int T = MODE == 1 ? level + 1 : n;      // tokens per sample whose outputs are needed (BOS + known bits)
    const int S = 64 / T;                          // samples per tile
    const int rows = S * T;
    const double scale = 1.0 / sqrt((double)hd);
    const int64_t ntiles = (B + S - 1) / S;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * S;
        __syncthreads();
        if (tid < S) s_idx[tid] = base + tid < B ? (uint64_t)idx_in[base + tid] : 0ull;
        __syncthreads();
        // ---- embedding: token (BOS = 2, then the bits) + position -------------------------------------------------------
        for (int e = tid; e < 64 * 64; e += MD_THREADS) {
            const int k = e >> 6, r = e & 63;
            double v = 0.0;
            if (r < rows) {
                const int s = r / T, t = r - s * T;
                const int tok = t == 0 ? 2 : (int)((s_idx[s] >> (t - 1)) & 1ull);
                v = __ldg(P.tok_emb + tok * TF_D + k) + __ldg(P.pos_emb + t * TF_D + k);
            }
            X[k * MD_S + r] = v;
        }
        for (int l = 0; l < P.depth; ++l) {
            // ---- q, k, v projections -----------------------------------------------------------------------------------
            double *dst[3] = {Qb, Kb, Vb};
            for (int part = 0; part < 3; ++part) {
                __syncthreads();
                load_weights_t(wt, P.in_proj_w[l], part * TF_D, TF_D, TF_D);
                __syncthreads();
                double acc[4][4];
                gemm_tile(X, wt, TF_D, tx, ty, acc, dst[part]);
                store_tile(dst[part], acc, P.in_proj_b[l] ? P.in_proj_b[l] + part * TF_D : nullptr, tx, ty);
            }
            __syncthreads();
            // ---- causal attention, one thread per (row, head); the result overwrites the thread's own query slice ------
            for (int pair = tid; pair < rows * H; pair += MD_THREADS) {
                const int r = pair / H, h = pair - r * H;
                const int s = r / T, t = r - s * T;
                const int r0 = s * T;
                double q[16];  // head_dim <= 16 is handled in registers; larger heads loop twice below
                for (int c0 = 0; c0 < hd; c0 += 16) {
                    const int cw = min(16, hd - c0);
                    for (int d = 0; d < cw; ++d) q[d] = Qb[(h * hd + c0 + d) * MD_S + r] * scale;
                    if (c0 == 0 && hd <= 16) break;
                }
                // pass 1: maximum score
                double mx = -INFINITY;
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0;
                    for (int d = 0; d < hd; ++d) sc += (hd <= 16 ? q[d] : Qb[(h * hd + d) * MD_S + r] * scale) * Kb[(h * hd + d) * MD_S + r0 + tp];
                    mx = fmax(mx, sc);
                }
                // pass 2: exp-weighted sum of the values
                double den = 0.0, o[16];
                for (int c0 = 0; c0 < hd; c0 += 16) {
                    const int cw = min(16, hd - c0);
                    for (int d = 0; d < cw; ++d) o[d] = 0.0;
                    den = 0.0;
                    for (int tp = 0; tp <= t; ++tp) {
                        double sc = 0.0;
                        for (int d = 0; d < hd; ++d) sc += (hd <= 16 ? q[d] : Qb[(h * hd + d) * MD_S + r] * scale) * Kb[(h * hd + d) * MD_S + r0 + tp];
                        const double p = exp(sc - mx);
                        den += p;
                        for (int d = 0; d < cw; ++d) o[d] += p * Vb[(h * hd + c0 + d) * MD_S + r0 + tp];
                    }
                    // queries of this slice are no longer needed once every slice has its scores: with hd <= 16 there is
                    // a single slice, otherwise the slices are written to Vb-independent storage after the loop
                    if (hd <= 16)
                        for (int d = 0; d < cw; ++d) Qb[(h * hd + c0 + d) * MD_S + r] = o[d] / den;
                    else
                        for (int d = 0; d < cw; ++d) wt[(h * hd + c0 + d) * MD_S + r] = o[d] / den;
                }
            }
            __syncthreads();
            if (hd > 16) {  // copy the attention output back (wt is rewritten by the next weight load)
                for (int e = tid; e < 64 * 64; e += MD_THREADS) Qb[(e >> 6) * MD_S + (e & 63)] = wt[(e >> 6) * MD_S + (e & 63)];
                __syncthreads();
            }
            // ---- output projection + residual + LayerNorm 1 ---------------------------------------------------------------
            load_weights_t(wt, P.out_proj_w[l], 0, TF_D, TF_D);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(Qb, wt, TF_D, tx, ty, acc, Kb);
                residual_layer_norm(X, acc, P.out_proj_b[l], P.ln1_w[l], P.ln1_b[l], P.ln_eps, tx, ty);
            }
            __syncthreads();
            // ---- feed-forward (dim -> dim, ReLU, dim -> dim) + residual + LayerNorm 2 -----------------------------------------
            load_weights_t(wt, P.lin1_w[l], 0, TF_D, TF_D);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(X, wt, TF_D, tx, ty, acc, Kb);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = tx + 16 * jj;
                    const double b = P.lin1_b[l] ? __ldg(P.lin1_b[l] + j) : 0.0;
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) Kb[j * MD_S + ty * 4 + ss] = fmax(acc[ss][jj] + b, 0.0);
                }
            }
            __syncthreads();
            load_weights_t(wt, P.lin2_w[l], 0, TF_D, TF_D);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(Kb, wt, TF_D, tx, ty, acc, Vb);
                residual_layer_norm(X, acc, P.lin2_b[l], P.ln2_w[l], P.ln2_b[l], P.ln_eps, tx, ty);
            }
            __syncthreads();
        }
        // ---- decoder: 4 numbers per token --------------------------------------------------------------------------------
        {
            const int r = tid >> 2, c = tid & 3;
            double acc = __ldg(P.dec_b + c);
            for (int k = 0; k < TF_D; ++k) acc = fma(X[k * MD_S + r], __ldg(P.dec_w + c * TF_D + k), acc);
            s_dec[r * 4 + c] = acc;
        }
        __syncthreads();
        // ---- masks, normalisation, gather (one thread per sample) ------------------------------------------------------------
        if (tid < S && base + tid < B) {
            const uint64_t x = s_idx[tid];
            const int t_lo = MODE == 1 ? level : 0, t_hi = MODE == 1 ? level + 1 : n;
            double re = 0.0, im = 0.0;
            bool dead = false;
            for (int t = t_lo; t < t_hi; ++t) {
                const double *o = s_dec + (tid * T + t) * 4;  // (re0, im0, re1, im1)
                const uint64_t prefix = t == 0 ? 0ull : (x & ((1ull << t) - 1ull));
                const long long mi = memo_index_of(P.sym_num, P.sym, prefix);
                const uint64_t mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)t * P.memo_size + mi) : 0ull;
                const bool a0 = mw & 1ull, a1 = (mw >> 1) & 1ull;
                const double z0 = a0 ? o[0] : -INFINITY, z1 = a1 ? o[2] : -INFINITY;
                const double mx = fmax(z0, z1);
                const double L = mx + 0.5 * log((a0 ? exp(2.0 * (z0 - mx)) : 0.0) + (a1 ? exp(2.0 * (z1 - mx)) : 0.0));
                if (MODE == 1) {
                    cond_out[(size_t)(base + tid) * 2 + 0] = a0 ? z0 - L : -INFINITY;
                    cond_out[(size_t)(base + tid) * 2 + 1] = a1 ? z1 - L : -INFINITY;
                } else {
                    const int bit = (int)((x >> t) & 1ull);
                    if (bit ? a1 : a0) {
                        re += (bit ? z1 : z0) - L;
                        im += bit ? o[3] : o[1];
                    } else {
                        dead = true;
                    }
                }
            }
            if (MODE == 0) log_psi[base + tid] = dead ? make_double2(-INFINITY, 0.0) : make_double2(re, im);
        }
    }
}

}  // namespace anqs

using namespace anqs;

static int tf_check(const anqs_transformer_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->dim == TF_D, "model dimension must be 64");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] encoder layers");
    ANQS_REQUIRE(P->head_num == 1 || P->head_num == 2 || P->head_num == 4 || P->head_num == 8 || P->head_num == 16,
                 "head_num must divide 64 and be at most 16");
    ANQS_REQUIRE(P->sym_num >= 0 && P->sym_num <= 8, "at most 8 symmetries");
    ANQS_REQUIRE(P->tok_emb && P->pos_emb && P->dec_w && P->dec_b, "null embedding / decoder pointer");
    for (int l = 0; l < P->depth; ++l)
        ANQS_REQUIRE(P->in_proj_w[l] && P->out_proj_w[l] && P->lin1_w[l] && P->lin2_w[l] && P->ln1_w[l] && P->ln1_b[l] && P->ln2_w[l] &&
                         P->ln2_b[l], "null layer weight pointer");
    ANQS_REQUIRE(P->cont_mask && P->memo_size >= 1, "null continuation-mask table");
    return 0;
}

template <int MODE>
static int tf_launch(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, int level, double *d_log_psi, double *d_cond,
                     void *stream) {
    auto kern = transformer_forward_kernel<MODE>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM));
    const int T = MODE == 1 ? level + 1 : desc->qubit_num;
    const int S = 64 / T;
    const int64_t ntiles = (n + S - 1) / S;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device());
    kern<<<grid, MD_THREADS, TF_SMEM, (cudaStream_t)stream>>>(*desc, d_idx, n, level, (double2 *)d_log_psi, d_cond);
    ANQS_LAUNCH_CHECK();
    return 0;
}

extern "C" {

int anqs_transformer_log_psi(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, void *stream) {
    if (tf_check(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_log_psi, "null pointer");
    return tf_launch<0>(desc, d_idx, n, 0, d_log_psi, nullptr, stream);
}

int anqs_transformer_cond_log_abs(const anqs_transformer_desc_t *desc, int qubit_idx, const int64_t *d_prefix, int64_t n,
                                  double *d_cond, void *stream) {
    if (tf_check(desc)) return 1;
    ANQS_REQUIRE(qubit_idx >= 0 && qubit_idx < desc->qubit_num, "qubit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_prefix && d_cond, "null pointer");
    return tf_launch<1>(desc, d_prefix, n, qubit_idx, nullptr, d_cond, stream);
}

}  // extern "C"
