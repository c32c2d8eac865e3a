// Kernel family 3, NADE mode on the tensor cores: the per-qudit MLP pairs of k3_nade.cu (reference ANQS:410-428, LAP:24-42,
// 63-134) with every layer a tcgen05.mma kind::tf32 GEMM (fp32 accumulators in TMEM) and the epilogue - bias, tanh, residual,
// mean over the qudit's own outcomes, continuation mask, 0.5 logsumexp(2 z), gather - in the fp32 registers of the thread that
// owns the sample.  Inference only; agreement with the fp64 kernel is a stated tolerance (tests/test_gpu_nade.py).
//
// One CTA = 128 threads = 128 samples (thread r = TMEM lane r), two CTAs per SM.  For every (sub-network, qudit) the packed
// block [W1 | W2.. | W_out | biases] (each matrix 64 x 64 in the MMA's K-major operand layout; the first layer zero-padded to
// 64 inputs, the output layer zero-padded to 64 rows) is staged by bulk TMA, then depth + 1 MMA groups of 128 x 64 x 64 run
// against the activation operand the threads rewrite in place.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace anqs {

constexpr int NTC_THREADS = 128;
constexpr uint32_t NTC_MAT = 64 * 64 * 4;

__host__ __device__ inline uint32_t ntc_block_bytes(int depth) { return (uint32_t)(depth + 1) * (NTC_MAT + 64 * 4); }

// packed[(net * Q + q)] = [W_0 .. W_depth (canonical layout)][b_0 .. b_depth (64 floats each)]
__global__ void nade_tc_pack_kernel(const anqs_nade_desc_t P, unsigned char *out) {
    const int Q = P.qudit_num, depth = P.depth;
    const uint32_t blk = ntc_block_bytes(depth);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int nq = 0; nq < 2 * Q; ++nq) {
        const int q = nq % Q;
        const double *const *tab = P.ptrs + (size_t)nq * (depth + 1) * 2;
        const int start = P.qudit_starts[q], D = 1 << (P.qudit_starts[q + 1] - start);
        const int K0 = start == 0 ? 1 : start;
        unsigned char *base = out + (size_t)nq * blk;
        for (int l = 0; l <= depth; ++l) {
            const double *W = tab[2 * l], *b = tab[2 * l + 1];
            const int rows = l == depth ? D : 64, K = l == 0 ? K0 : 64;
            for (int64_t e = t0; e < 64 * 64; e += stride) {
                const int j = (int)(e >> 6), k = (int)(e & 63);
                // the first qudit's network sees one constant-zero input (LAP:26): its first layer contributes the bias only
                const float v = (j < rows && k < K && !(l == 0 && start == 0)) ? (float)W[(size_t)j * K + k] : 0.0f;
                *reinterpret_cast<float *>(base + (uint32_t)l * NTC_MAT + canon_off((uint32_t)j, (uint32_t)k, 64)) = v;
            }
            float *bv = reinterpret_cast<float *>(base + (uint32_t)(depth + 1) * NTC_MAT) + 64 * l;
            for (int64_t e = t0; e < 64; e += stride) bv[e] = (b && e < rows) ? (float)b[e] : 0.0f;
        }
    }
}

template <int MODE>  // 0: log psi, 1: conditional log|psi| of qudit level_q
__global__ void __launch_bounds__(NTC_THREADS, 2)
nade_tc_kernel(const anqs_nade_desc_t P, const unsigned char *__restrict__ packed, const int64_t *__restrict__ idx_in, int64_t B,
               int level_q, double2 *__restrict__ log_psi, double *__restrict__ cond_out) {
    extern __shared__ __align__(1024) unsigned char ntc_smem[];
    unsigned char *A = ntc_smem;                       // 32 KB activation operand
    unsigned char *W = A + 128 * 64 * 4;               // one (sub-network, qudit) block
    __shared__ uint64_t bar_w, bar_m;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, row = tid;
    const int Q = P.qudit_num, DM = P.max_qudit_dim, depth = P.depth;
    const uint32_t blk = ntc_block_bytes(depth);
    const float *bias = reinterpret_cast<const float *>(W + (uint32_t)(depth + 1) * NTC_MAT);
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_m, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t p_w = 0, p_m = 0;

    const int nets = MODE == 1 ? 1 : 2;
    const int q_lo = MODE == 1 ? level_q : 0, q_hi = MODE == 1 ? level_q + 1 : Q;
    const int64_t ntiles = (B + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t s = tile * 128 + row;
        const uint64_t x = s < B ? (uint64_t)idx_in[s] : 0ull;
        double acc_re = 0.0, acc_im = 0.0;
        bool dead = false;
        for (int net = 0; net < nets; ++net) {
            for (int q = q_lo; q < q_hi; ++q) {
                const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start, D = 1 << bits;
                __syncthreads();  // the previous block's last MMA and the readers of its biases are done
                if (tid == 0) {
                    mbar_arrive_expect_tx(&bar_w, blk);
                    const unsigned char *src = packed + (size_t)(net * Q + q) * blk;
                    for (uint32_t off = 0; off < blk; off += 32768u) bulk_copy_g2s(W + off, src + off, min(32768u, blk - off), &bar_w);
                }
                // input encoding: 1 - 2 bit for the qubits before the qudit, 0 beyond (and everywhere for the first qudit)
#pragma unroll
                for (int kb = 0; kb < 64; kb += 4) {
                    float4 v;
                    float *vp = &v.x;
#pragma unroll
                    for (int j = 0; j < 4; ++j) vp[j] = (kb + j) < start ? 1.0f - 2.0f * (float)((x >> (kb + j)) & 1ull) : 0.0f;
                    *reinterpret_cast<float4 *>(A + canon_off((uint32_t)row, (uint32_t)kb, 64)) = v;
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                mbar_wait(&bar_w, p_w);
                p_w ^= 1u;
                for (int l = 0; l <= depth; ++l) {
                    if (tid == 0) {
                        tc_fence_after();
                        issue_gemm(tmem, smem_u32(A), smem_u32(W) + (uint32_t)l * NTC_MAT, 64);
                        umma_commit(&bar_m);
                    }
                    mbar_wait(&bar_m, p_m);
                    p_m ^= 1u;
                    tc_fence_after();
                    float v[64];
                    tmem_ld64(my_tmem, v);
                    const float *bl = bias + 64 * l;
                    if (l < depth) {
                        // hidden layer: tanh(v + b (+ previous activation on the residual layers, MLP:237-239)), rewritten in place
#pragma unroll
                        for (int kb = 0; kb < 64; kb += 4) {
                            float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (P.use_res && l > 0) res = *reinterpret_cast<const float4 *>(A + canon_off((uint32_t)row, (uint32_t)kb, 64));
                            const float *rp = &res.x;
                            float4 o;
                            float *op = &o.x;
#pragma unroll
                            for (int j = 0; j < 4; ++j) op[j] = tanh_fast(v[kb + j] + bl[kb + j] + rp[j]);
                            *reinterpret_cast<float4 *>(A + canon_off((uint32_t)row, (uint32_t)kb, 64)) = o;
                        }
                        fence_proxy_async();
                        tc_fence_before();
                        __syncthreads();
                    } else {
                        const int chosen = (int)((x >> start) & ((1ull << bits) - 1ull));
                        if (net == 0) {
                            uint64_t mw;
                            if (P.du[q]) {
                                mw = D >= 64 ? ~0ull : ((1ull << D) - 1ull);
                            } else {
                                const uint64_t prefix = start == 0 ? 0ull : (x & ((1ull << start) - 1ull));
                                long long mi = 0;
                                for (int sy = 0; sy < P.sym_num; ++sy) {
                                    const int64_t *d = P.sym[sy];
                                    long long ev;
                                    if (d[0] == 0) ev = d[7] + __popcll(prefix & (uint64_t)d[1]) - __popcll(prefix & (uint64_t)d[2]);
                                    else ev = (__popcll(prefix & (uint64_t)d[1]) & 1) ? -d[7] : d[7];
                                    const long long num = ev * d[3] + d[4], qd = num / d[5], rm = num % d[5];
                                    mi += ((rm != 0 && ((rm < 0) != (d[5] < 0))) ? qd - 1 : qd) * d[6];
                                }
                                mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)q * P.memo_size + mi) : 0ull;
                            }
                            const uint32_t mlo = (uint32_t)mw, mhi = (uint32_t)(mw >> 32);
                            float sum = 0.f;
#pragma unroll
                            for (int d = 0; d < 64; ++d) {
                                v[d] += bl[d];
                                sum += d < D ? v[d] : 0.f;
                            }
                            const float mean = P.subtract_mean ? sum / (float)D : 0.f;  // over the qudit's own outcomes (LAP:118-119)
                            float mx = -INFINITY;
#pragma unroll
                            for (int d = 0; d < 64; ++d) {
                                const bool ok = d < D && ((d < 32 ? mlo >> d : mhi >> (d - 32)) & 1u) != 0u;
                                v[d] = ok ? v[d] - mean : -INFINITY;
                                mx = fmaxf(mx, v[d]);
                            }
                            const bool any = mw != 0ull && mx > -INFINITY;
                            const float mxs = any ? mx : 0.f;
                            float se = 0.f;
#pragma unroll
                            for (int d = 0; d < 64; ++d) se += __expf(2.0f * (v[d] - mxs));
                            const float Ln = mxs + 0.5f * __logf(se);
                            if (MODE == 1) {
                                if (s < B) {
#pragma unroll
                                    for (int d = 0; d < 64; ++d)
                                        if (d < DM) cond_out[(size_t)s * DM + d] = (any && v[d] > -INFINITY) ? (double)(v[d] - Ln) : -INFINITY;
                                }
                            } else {
                                float pick = -INFINITY;
#pragma unroll
                                for (int d = 0; d < 64; ++d)
                                    if (d == chosen) pick = v[d];
                                if (any && pick > -INFINITY) acc_re += (double)(pick - Ln);
                                else dead = true;
                            }
                        } else {
                            float pick = 0.f;
#pragma unroll
                            for (int d = 0; d < 64; ++d)
                                if (d == chosen) pick = v[d] + bl[d];
                            acc_im += (double)pick;
                        }
                        tc_fence_before();
                    }
                }
            }
        }
        if (MODE == 0 && s < B) log_psi[s] = dead ? make_double2(-INFINITY, 0.0) : make_double2(acc_re, 3.14159265358979323846 * acc_im);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace anqs

using namespace anqs;

static int ntc_check(const anqs_nade_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->qudit_num >= 1 && P->qudit_num <= 64, "qudit_num must be in [1, 64]");
    ANQS_REQUIRE(P->max_qudit_dim >= 2 && P->max_qudit_dim <= 64, "max_qudit_dim must be in [2, 64]");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] hidden layers");
    ANQS_REQUIRE(P->width == 64, "hidden width must be 64 (the reference default)");
    ANQS_REQUIRE(P->ptrs && P->cont_mask && P->memo_size >= 1, "null pointer table / continuation-mask table");
    return 0;
}

template <int MODE>
static int ntc_launch(const anqs_nade_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n, int level_q, double *d_log_psi,
                      double *d_cond, void *stream) {
    auto kern = nade_tc_kernel<MODE>;
    const size_t smem = (size_t)128 * 64 * 4 + ntc_block_bytes(desc->depth);
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (n + 127) / 128;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device() * 2);
    kern<<<grid, NTC_THREADS, smem, (cudaStream_t)stream>>>(*desc, (const unsigned char *)d_packed, d_idx, n, level_q, (double2 *)d_log_psi, d_cond);
    ANQS_LAUNCH_CHECK();
    return 0;
}

extern "C" {

size_t anqs_nade_tc_packed_bytes(const anqs_nade_desc_t *desc) {
    if (!desc) return 0;
    return (size_t)2 * desc->qudit_num * ntc_block_bytes(desc->depth);
}

int anqs_nade_tc_pack(const anqs_nade_desc_t *desc, void *d_packed, void *stream) {
    if (ntc_check(desc)) return 1;
    ANQS_REQUIRE(d_packed && ((uintptr_t)d_packed & 127) == 0, "packed buffer must be non-null and 128-byte aligned");
    nade_tc_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*desc, (unsigned char *)d_packed);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_nade_log_psi_tc(const anqs_nade_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n, double *d_log_psi,
                         void *stream) {
    if (ntc_check(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_idx && d_log_psi, "null pointer");
    return ntc_launch<0>(desc, d_packed, d_idx, n, 0, d_log_psi, nullptr, stream);
}

int anqs_nade_cond_log_abs_tc(const anqs_nade_desc_t *desc, const void *d_packed, int qudit_idx, const int64_t *d_prefix, int64_t n,
                              double *d_cond, void *stream) {
    if (ntc_check(desc)) return 1;
    ANQS_REQUIRE(qudit_idx >= 0 && qudit_idx < desc->qudit_num, "qudit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_prefix && d_cond, "null pointer");
    return ntc_launch<1>(desc, d_packed, d_prefix, n, qudit_idx, nullptr, d_cond, stream);
}

}  // extern "C"
