// Kernel family 3: MADE conditional log-amplitudes / amplitudes with the symmetry masks fused in.
//
//   made_forward_kernel<LOGPSI>  packed configurations -> log psi = sum_q [ masked, normalised conditional
//                                log|psi| at the chosen outcome ] + i*pi*sum_q phase   (reference ANQS:407-485 with
//                                LAP:63-163, MLP:217-246, QG:199-213, ANQS:392-405)
//   made_forward_kernel<COND>    prefixes -> normalised conditional log|psi| of ONE qudit (what the samplers call per
//                                level: ANQS:615-620 / 718-723 -> LAP:105-163)
//
// fp64 parity path (the reference network is float64, constants.py:4): every GEMM is CUDA-core DFMA on a
// 64-sample x 64-output tile, 4x4 accumulators per thread, operands in shared memory.  Nothing but the packed
// int64 configuration is read per sample and nothing but log psi (16 B) is written: the bit unpacking (HS:121-132),
// the 1-2b input encoding (MLP:205-215), the rolling quantum numbers (MSK:156-167), the continuation masks
// (QG:199-213), the pre-mask mean subtraction (ANQS:338-340), the masked logsumexp normalisation (ANQS:392-405)
// and the gather of the chosen outcome (ANQS:450-454) all happen in registers / shared memory.  The phase network's
// last layer is evaluated only at the chosen outcome of every qudit (Q rows instead of Q*D).
#include <algorithm>

#include "common.cuh"
#include "made_common.cuh"

namespace anqs {

constexpr size_t MD_SMEM = (size_t)3 * 64 * MD_S * sizeof(double) + 3 * 64 * sizeof(uint64_t) + 4 * 64 * sizeof(double);

constexpr int MADE_LOGPSI = 0, MADE_COND = 1;

template <int MODE>
__global__ void __launch_bounds__(MD_THREADS, 2)
made_forward_kernel(const anqs_made_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, int level_q,
                    double2 *__restrict__ log_psi, double *__restrict__ cond_out, double *__restrict__ save_h,
                    double *__restrict__ save_p) {
    extern __shared__ __align__(16) unsigned char md_smem[];
    double *act0 = reinterpret_cast<double *>(md_smem);
    double *act1 = act0 + 64 * MD_S;
    double *wt = act1 + 64 * MD_S;
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(wt + 64 * MD_S);
    uint64_t *s_mask = s_idx + 64;
    uint64_t *s_pref = s_mask + 64;  // (unused slot kept for alignment)
    double *s_im = reinterpret_cast<double *>(s_pref + 64);  // [4][64] phase partial sums
    (void)s_pref;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = P.qubit_num, Q = P.qudit_num, DM = P.max_qudit_dim, depth = P.depth;
    const int known = MODE == MADE_COND ? P.qudit_starts[level_q] : n;
    const int64_t ntiles = (B + MD_TB - 1) / MD_TB;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * MD_TB;
        __syncthreads();
        if (tid < 64) {
            uint64_t x = base + tid < B ? (uint64_t)idx_in[base + tid] : 0ull;
            if (MODE == MADE_COND) x = known >= 64 ? x : (x & ((1ull << known) - 1ull));
            s_idx[tid] = x;
        }
        double out_re[4] = {0.0, 0.0, 0.0, 0.0};
        const int nets = MODE == MADE_COND ? 1 : 2;
        for (int net = 0; net < nets; ++net) {
            const double *const *Ws = net == 0 ? P.w_abs : P.w_phase;
            const double *const *bs = net == 0 ? P.b_abs : P.b_phase;
            __syncthreads();
            double wv[16];   // weight tile in flight: fetched one product ahead, committed to shared memory when `wt` is free
            fetch_weights_t(wv, Ws[0], 0, MD_W, n);
            const int q_lo = MODE == MADE_COND ? level_q : 0, q_hi = MODE == MADE_COND ? level_q + 1 : Q;
            // input encoding: 1 - 2*bit for known positions, 0 beyond the prefix (MLP:205-225)
            for (int e = tid; e < n * 64; e += MD_THREADS) {
                int k = e >> 6, s = e & 63;
                double v = 1.0 - 2.0 * (double)((s_idx[s] >> k) & 1ull);
                act0[k * MD_S + s] = k < known ? v : 0.0;
            }
            double *cur = act0, *nxt = act1;
            for (int l = 0; l < depth; ++l) {
                const int K = l == 0 ? n : MD_W;
                commit_weights_t(wt, wv, K);
                if (l + 1 < depth) fetch_weights_t(wv, Ws[l + 1], 0, MD_W, MD_W);
                else if (net == 0) fetch_weights_t(wv, Ws[depth], q_lo * DM, DM, MD_W);
                __syncthreads();
                double acc[4][4];
                gemm_tile(cur, wt, K, tx, ty, acc, nxt);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = tx + 16 * jj;
                    const double bias = bs[l] ? __ldg(bs[l] + j) : 0.0;
#pragma unroll
                    for (int ss = 0; ss < 4; ++ss) {
                        double v = acc[ss][jj] + bias;
                        if (P.use_res && l > 0) v += cur[j * MD_S + ty * 4 + ss];  // MLP:237-239
                        v = tanh(v);
                        nxt[j * MD_S + ty * 4 + ss] = v;
                        if (save_h && base + ty * 4 + ss < B)
                            save_h[(((size_t)net * depth + l) * (size_t)B + (size_t)(base + ty * 4 + ss)) * MD_W + j] = v;
                    }
                }
                __syncthreads();
                double *t = cur;
                cur = nxt;
                nxt = t;
            }
            // cur = last hidden activations [64][samples]
            if (net == 0) {
                for (int q = q_lo; q < q_hi; ++q) {
                    commit_weights_t(wt, wv, MD_W);
                    if (q + 1 < q_hi) fetch_weights_t(wv, Ws[depth], (q + 1) * DM, DM, MD_W);
                    if (tid < 64) {
                        const int start = P.qudit_starts[q];
                        const uint64_t x = s_idx[tid];
                        const uint64_t prefix = start == 0 ? 0ull : (x & ((1ull << start) - 1ull));
                        uint64_t mw;
                        if (P.du[q]) {
                            // Unmasked qudit.  The reference is not consistent about how wide "unmasked" is in MADE mode, and both
                            // widths are reproduced: log psi / amplitudes unmask all max_qudit_dim output columns (the mask is
                            // padded first and replaced by ones afterwards, ANQS:434-441), the samplers' conditionals only the
                            // qudit's own 2^bits outcomes (ones first, padding afterwards, ANQS:605-613, 708-716).
                            const int dq = MODE == MADE_COND ? (1 << (P.qudit_starts[q + 1] - start)) : DM;
                            mw = dq >= 64 ? ~0ull : ((1ull << dq) - 1ull);
                        } else {
                            long long mi = memo_index(P, prefix);
                            mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)q * P.memo_size + mi) : 0ull;
                        }
                        s_mask[tid] = mw;
                    }
                    __syncthreads();
                    double acc[4][4];
                    gemm_tile(cur, wt, MD_W, tx, ty, acc, nxt);
                    const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
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
                            double bias = (bs[depth] && d < DM) ? __ldg(bs[depth] + q * DM + d) : 0.0;
                            z[jj] = acc[ss][jj] + bias;
                            sum += d < DM ? z[jj] : 0.0;
                        }
                        if (P.subtract_mean) {  // over ALL max_qudit_dim entries, before masking (ANQS:338-340)
                            const double mean = row_sum16(sum) / (double)DM;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) z[jj] -= mean;
                        }
                        double mx = -INFINITY;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int d = tx + 16 * jj;
                            if (d < DM && ((mw >> d) & 1ull)) mx = fmax(mx, z[jj]);
                        }
                        mx = row_max16(mx);
                        double se = 0.0, ez[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int d = tx + 16 * jj;
                            ez[jj] = (d < DM && ((mw >> d) & 1ull)) ? exp(2.0 * (z[jj] - mx)) : 0.0;
                            se += ez[jj];
                        }
                        se = row_sum16(se);
                        const double L = mx + 0.5 * log(se);  // 0.5 * logsumexp(2 z) over the allowed outcomes
                        const bool any = mw != 0ull && mx > -INFINITY;
                        if (MODE == MADE_COND) {
                            if (base + s < B) {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int d = tx + 16 * jj;
                                    if (d < DM)
                                        cond_out[(size_t)(base + s) * DM + d] =
                                            (any && ((mw >> d) & 1ull)) ? z[jj] - L : -INFINITY;
                                }
                            }
                        } else {
                            double pick = 0.0;
                            const double inv_se = 1.0 / se;   // p_d = exp(2 (z_d - L)) = exp(2 (z_d - max)) / sum: no second exponential
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int d = tx + 16 * jj;
                                const bool allowed = any && d < DM && ((mw >> d) & 1ull);
                                if (d == chosen) pick = allowed ? z[jj] - L : -INFINITY;
                                if (save_p && d < DM && base + s < B)
                                    save_p[((size_t)(base + s) * Q + q) * DM + d] = allowed ? ez[jj] * inv_se : 0.0;
                            }
                            out_re[ss] += row_sum16(pick);
                        }
                    }
                    __syncthreads();
                }
            } else {
                // phase network: only the row of the chosen outcome of every qudit (LAP:97, ANQS:450-454)
                const int s = tid & 63, g = tid >> 6;
                const uint64_t x = s_idx[s];
                double part = 0.0;
                for (int q = g; q < Q; q += 4) {
                    const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
                    const int row = q * DM + (int)((x >> start) & ((1ull << bits) - 1ull));
                    // the row is 512 contiguous bytes: 16-byte loads (every 32-byte sector is used up by two instructions), eight
                    // of them issued before the first use, four partial sums instead of one dependent chain
                    const double2 *w = reinterpret_cast<const double2 *>(Ws[depth] + (size_t)row * MD_W);
                    double d0 = bs[depth] ? __ldg(bs[depth] + row) : 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                    for (int k0 = 0; k0 < MD_W; k0 += 16) {
                        double2 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __ldg(w + (k0 >> 1) + i);
#pragma unroll
                        for (int i = 0; i < 8; i += 2) {
                            d0 = fma(v[i].x, cur[(k0 + 2 * i) * MD_S + s], d0);
                            d1 = fma(v[i].y, cur[(k0 + 2 * i + 1) * MD_S + s], d1);
                            d2 = fma(v[i + 1].x, cur[(k0 + 2 * i + 2) * MD_S + s], d2);
                            d3 = fma(v[i + 1].y, cur[(k0 + 2 * i + 3) * MD_S + s], d3);
                        }
                    }
                    part += (d0 + d1) + (d2 + d3);
                }
                s_im[g * 64 + s] = part;
                __syncthreads();
            }
        }
        if (MODE == MADE_LOGPSI) {
            // every lane of a 16-lane row group holds out_re; one of them publishes it
            double *s_re = wt;  // weights are dead here
            __syncthreads();
            if (tx == 0) {
#pragma unroll
                for (int ss = 0; ss < 4; ++ss) s_re[ty * 4 + ss] = out_re[ss];
            }
            __syncthreads();
            if (tid < 64 && base + tid < B) {
                const double im = s_im[tid] + s_im[64 + tid] + s_im[128 + tid] + s_im[192 + tid];
                const double re = s_re[tid];
                // ANQS:399-401: an unphysical configuration ends as (-inf, 0)
                log_psi[base + tid] = make_double2(re, re == -INFINITY ? 0.0 : 3.14159265358979323846 * im);
            }
        }
    }
}

}  // namespace anqs

using namespace anqs;

static int check_desc(const anqs_made_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->qudit_num >= 1 && P->qudit_num <= 64, "qudit_num must be in [1, 64]");
    ANQS_REQUIRE(P->max_qudit_dim >= 2 && P->max_qudit_dim <= 64, "max_qudit_dim must be in [2, 64]");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] hidden layers");
    ANQS_REQUIRE(P->width == MD_W, "hidden width must be 64 (the reference default)");
    ANQS_REQUIRE(P->sym_num >= 0 && P->sym_num <= 8, "at most 8 symmetries");
    ANQS_REQUIRE(P->qudit_starts[0] == 0 && P->qudit_starts[P->qudit_num] == P->qubit_num, "qudit_starts must span the qubits");
    for (int q = 0; q < P->qudit_num; ++q) {
        int bits = P->qudit_starts[q + 1] - P->qudit_starts[q];
        ANQS_REQUIRE(bits >= 1 && (1 << bits) <= P->max_qudit_dim, "qudit wider than max_qudit_dim");
    }
    for (int l = 0; l <= P->depth; ++l) ANQS_REQUIRE(P->w_abs[l] && P->w_phase[l], "null weight pointer");
    ANQS_REQUIRE(P->cont_mask && P->memo_size >= 1, "null continuation-mask table");
    return 0;
}

extern "C" {

int anqs_made_log_psi(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi,
                      double *d_save_h, double *d_save_p, void *stream) {
    if (check_desc(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_log_psi, "null pointer");
    auto kern = made_forward_kernel<MADE_LOGPSI>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM));
    int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, MD_THREADS, MD_SMEM, (cudaStream_t)stream>>>(*desc, d_idx, n, 0, (double2 *)d_log_psi, nullptr, d_save_h,
                                                              d_save_p);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_made_cond_log_abs(const anqs_made_desc_t *desc, int qudit_idx, const int64_t *d_prefix, int64_t n,
                           double *d_cond, void *stream) {
    if (check_desc(desc)) return 1;
    ANQS_REQUIRE(qudit_idx >= 0 && qudit_idx < desc->qudit_num, "qudit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_prefix && d_cond, "null pointer");
    auto kern = made_forward_kernel<MADE_COND>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM));
    int64_t ntiles = (n + MD_TB - 1) / MD_TB;
    int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, MD_THREADS, MD_SMEM, (cudaStream_t)stream>>>(*desc, d_prefix, n, qudit_idx, nullptr, d_cond, nullptr, nullptr);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
