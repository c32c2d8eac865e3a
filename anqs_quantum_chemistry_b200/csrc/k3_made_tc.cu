// Kernel family 3, tensor-core mode: MADE conditional log-amplitudes / amplitudes with every GEMM on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32: fp32 operands in shared memory, fp32 accumulators in TMEM) and the symmetry
// masks, mean subtraction, masked softmax normalisation and outcome gather fused into the TMEM read-back.
// Same inputs and outputs as made_forward_kernel (k3_made.cu, the fp64 parity mode; reference ANQS:309-485,
// LAP:63-163, MLP:217-246); agreement with it is a stated tolerance (tests/test_gpu_anqs.py), not 1e-10: tf32
// multiplies carry 10 mantissa bits.
//
// One CTA = 128 threads = 128 samples (one TMEM lane and one accumulator row per thread), persistent over tiles.
//   input / hidden activations : [128 x K] fp32 in shared memory, canonical no-swizzle K-major core-matrix layout
//                                (8 rows x 16 bytes per core matrix; LBO = 128 B between K-adjacent core matrices,
//                                SBO = K/4 * 128 B between 8-row groups), written by the thread that owns the row
//   weights                    : packed once per parameter update by made_tc_pack_kernel into the same layout
//                                ([out][in] nn.Linear weights are already N x K, K-major), staged per use by 1-D bulk
//                                TMA copies (cp.async.bulk + mbarrier); the output layer streams in chunks of two qudits
//                                (2 x 64 rows, 32 KB) through a double buffer
//   MMA                        : M = 128, N = 64, K = 8 per instruction, issued by thread 0, completion through
//                                tcgen05.commit -> mbarrier; the output-layer MMA of chunk c+1 runs under the epilogue of
//                                chunk c (two TMEM accumulator buffers)
//   epilogue                   : tcgen05.ld 32x32b.x64 gives a thread the 64 logits of its sample for one qudit:
//                                bias, pre-mask mean (ANQS:338-340), continuation mask (QG:199-213), 0.5*logsumexp(2z)
//                                (ANQS:392-405) and the chosen outcome (ANQS:450-454) are register work, no shuffles
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace anqs {

constexpr int TC_NT = 4;                        // 128-sample tiles per CTA
constexpr int TC_THREADS = 128 * TC_NT;
constexpr int TC_W = 64;                        // hidden width and outcomes per qudit
constexpr uint32_t TC_ACT_BYTES = 128 * 64 * 4;  // one activation tile
constexpr uint32_t TC_WH_BYTES = 64 * 64 * 4;    // one hidden-layer weight matrix / one qudit block of the output layer
constexpr uint32_t TC_TMEM_COLS = 512;

struct TcPacked {                 // offsets (bytes) into the packed buffer, per sub-network
    uint32_t w1[2], w2[2], w3[2]; // w3: qudit_num blocks of TC_WH_BYTES
    uint32_t b1[2], b2[2], b3[2]; // fp32 biases
    uint32_t k0pad;               // input width rounded up to a multiple of 8
    uint32_t total;
};

__host__ __device__ inline TcPacked tc_layout(int qubit_num, int qudit_num, int depth) {
    TcPacked L;
    L.k0pad = (uint32_t)((qubit_num + 7) / 8 * 8);
    uint32_t off = 0;
    for (int net = 0; net < 2; ++net) {
        L.w1[net] = off; off += 64 * L.k0pad * 4;
        L.w2[net] = off; off += (uint32_t)(depth - 1) * TC_WH_BYTES;
        L.w3[net] = off; off += (uint32_t)qudit_num * TC_WH_BYTES;
        L.b1[net] = off; off += 64 * 4;
        L.b2[net] = off; off += (uint32_t)(depth - 1) * 64 * 4;
        L.b3[net] = off; off += (uint32_t)qudit_num * 64 * 4;
        off = (off + 127) / 128 * 128;
    }
    L.total = off;
    return L;
}

__global__ void made_tc_pack_kernel(const anqs_made_desc_t P, TcPacked L, unsigned char *out) {
    const int Q = P.qudit_num, DM = P.max_qudit_dim, n = P.qubit_num, depth = P.depth;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int net = 0; net < 2; ++net) {
        const double *const *Ws = net == 0 ? P.w_abs : P.w_phase;
        const double *const *bs = net == 0 ? P.b_abs : P.b_phase;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 64 * (int64_t)L.k0pad; e += stride) {
            const uint32_t j = (uint32_t)(e / L.k0pad), k = (uint32_t)(e % L.k0pad);
            *reinterpret_cast<float *>(out + L.w1[net] + canon_off(j, k, L.k0pad)) = k < (uint32_t)n ? (float)Ws[0][(size_t)j * n + k] : 0.0f;
        }
        for (int l = 1; l < depth; ++l)
            for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 64 * 64; e += stride) {
                const uint32_t j = (uint32_t)(e >> 6), k = (uint32_t)(e & 63);
                *reinterpret_cast<float *>(out + L.w2[net] + (uint32_t)(l - 1) * TC_WH_BYTES + canon_off(j, k, 64)) = (float)Ws[l][(size_t)j * 64 + k];
            }
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)Q * 64 * 64; e += stride) {
            const uint32_t q = (uint32_t)(e >> 12), d = (uint32_t)((e >> 6) & 63), k = (uint32_t)(e & 63);
            const float v = d < (uint32_t)DM ? (float)Ws[depth][((size_t)q * DM + d) * 64 + k] : 0.0f;
            *reinterpret_cast<float *>(out + L.w3[net] + q * TC_WH_BYTES + canon_off(d, k, 64)) = v;
        }
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < 64; e += stride) {
            reinterpret_cast<float *>(out + L.b1[net])[e] = bs[0] ? (float)bs[0][e] : 0.0f;
            for (int l = 1; l < depth; ++l) reinterpret_cast<float *>(out + L.b2[net])[(l - 1) * 64 + e] = bs[l] ? (float)bs[l][e] : 0.0f;
        }
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)Q * 64; e += stride) {
            const int q = (int)(e >> 6), d = (int)(e & 63);
            reinterpret_cast<float *>(out + L.b3[net])[e] = (bs[depth] && d < DM) ? (float)bs[depth][q * DM + d] : 0.0f;
        }
    }
}

__device__ __forceinline__ long long tc_floor_div(long long a, long long b) {
    long long q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ long long tc_memo_index(const anqs_made_desc_t &P, uint64_t prefix) {
    long long idx = 0;
    for (int s = 0; s < P.sym_num; ++s) {
        const int64_t *d = P.sym[s];
        long long e;
        if (d[0] == 0)
            e = d[7] + __popcll(prefix & (uint64_t)d[1]) - __popcll(prefix & (uint64_t)d[2]);
        else
            e = (__popcll(prefix & (uint64_t)d[1]) & 1) ? -d[7] : d[7];
        idx += tc_floor_div(e * d[3] + d[4], d[5]) * d[6];
    }
    return idx;
}

// mode 0: log psi of whole configurations; mode 1: normalised conditional log|psi| of qudit level_q for prefixes.
// One CTA works on TC_NT tiles of 128 samples at once (4 warps per tile): every GEMM phase issues one MMA group per tile
// against the same staged weights, and all 4 * TC_NT warps run the epilogue in parallel.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
made_tc_kernel(const anqs_made_desc_t P, const TcPacked L, const unsigned char *__restrict__ packed,
               const int64_t *__restrict__ idx_in, int64_t B, int level_q, double2 *__restrict__ log_psi,
               double *__restrict__ cond_out) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *acts = tc_smem;                                  // TC_NT activation tiles, rewritten in place per layer
    unsigned char *wh = acts + TC_NT * TC_ACT_BYTES;                // hidden-layer weights
    unsigned char *wc = wh + TC_WH_BYTES;                           // 2 x one qudit block of the output layer
    float *s_bias = reinterpret_cast<float *>(wc + 2 * TC_WH_BYTES);  // [b1 64][b2 64*(depth-1)][b3 Q*64]
    __shared__ uint64_t bar_wh, bar_wc[2], bar_m[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int tl = tid >> 7, row = tid & 127;                       // this thread's tile and row
    unsigned char *act = acts + tl * TC_ACT_BYTES;
    const int n = P.qubit_num, Q = P.qudit_num, depth = P.depth;
    const uint32_t K0 = L.k0pad;
    if (tid == 0) {
        mbar_init(&bar_wh, 1);
        mbar_init(&bar_wc[0], 1);
        mbar_init(&bar_wc[1], 1);
        mbar_init(&bar_m[0], 1);
        mbar_init(&bar_m[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // TMEM: tile t owns columns [128 t, 128 t + 128): two 64-column accumulator buffers (buffer 0 also serves the hidden layers)
    const uint32_t my_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)tl * 128u;
    uint32_t p_wh = 0, p_wc[2] = {0, 0}, p_m[2] = {0, 0};

    const int nets = MODE == 1 ? 1 : 2;
    const int q_lo = MODE == 1 ? level_q : 0, q_hi = MODE == 1 ? level_q + 1 : Q;
    const int nchunks = q_hi - q_lo;                                // one qudit per chunk
    const int known = MODE == 1 ? P.qudit_starts[level_q] : n;
    const int64_t ngroups = (B + 128 * TC_NT - 1) / (128 * TC_NT);

    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t s = (grp * TC_NT + tl) * 128 + row;
        const int64_t left_tiles = (B - grp * TC_NT * 128 + 127) / 128;
        const int live_tiles = left_tiles < TC_NT ? (int)left_tiles : TC_NT;  // tiles of this group that hold samples
        uint64_t x = s < B ? (uint64_t)idx_in[s] : 0ull;
        if (MODE == 1) x = known >= 64 ? x : (x & ((1ull << known) - 1ull));
        double acc_re = 0.0, acc_im = 0.0;
        bool dead = false;  // an unphysical configuration: log|psi| = -inf (ANQS:399-401)

        for (int net = 0; net < nets; ++net) {
            // ---- stage weights: W1 now; the first two output blocks as early as possible ------------------------------
            __syncthreads();  // previous users of wh / wc / bias / act buffers are done
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_wh, 64 * K0 * 4);
                bulk_copy_g2s(wh, packed + L.w1[net], 64 * K0 * 4, &bar_wh);
                for (int c = 0; c < nchunks && c < 2; ++c) {
                    mbar_arrive_expect_tx(&bar_wc[c], TC_WH_BYTES);
                    bulk_copy_g2s(wc + c * TC_WH_BYTES, packed + L.w3[net] + (uint32_t)(q_lo + c) * TC_WH_BYTES, TC_WH_BYTES, &bar_wc[c]);
                }
            }
            for (int e = tid; e < 64 * depth + Q * 64; e += TC_THREADS)
                s_bias[e] = reinterpret_cast<const float *>(packed + L.b1[net])[e];  // b1 | b2 | b3 are contiguous
            // input encoding: 1 - 2 bit for the known positions, 0 beyond the prefix and in the padding (MLP:205-225)
            for (uint32_t kb = 0; kb < K0; kb += 4) {
                float4 v;
                float *vp = &v.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = (int)kb + j;
                    vp[j] = k < known ? 1.0f - 2.0f * (float)((x >> k) & 1ull) : 0.0f;
                }
                *reinterpret_cast<float4 *>(act + canon_off((uint32_t)row, kb, K0)) = v;
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- hidden layers (activations rewritten in place: the MMA has finished reading before the epilogue writes) --
            for (int l = 0; l < depth; ++l) {
                const uint32_t K = l == 0 ? K0 : 64u;
                if (tid == 0) {
                    mbar_wait(&bar_wh, p_wh);
                    tc_fence_after();
                    for (int t = 0; t < live_tiles; ++t)
                        issue_gemm(tmem + 128u * t, smem_u32(acts + t * TC_ACT_BYTES), smem_u32(wh), K);
                    umma_commit(&bar_m[0]);
                }
                p_wh ^= 1u;
                mbar_wait(&bar_m[0], p_m[0]);
                p_m[0] ^= 1u;
                tc_fence_after();
                if (tid == 0 && l + 1 < depth) {  // the MMA is done reading wh: fetch the next hidden layer's weights
                    mbar_arrive_expect_tx(&bar_wh, TC_WH_BYTES);
                    bulk_copy_g2s(wh, packed + L.w2[net] + (uint32_t)l * TC_WH_BYTES, TC_WH_BYTES, &bar_wh);
                }
                if (tl < live_tiles) {
                    float v[64];
                    tmem_ld64(my_tmem, v);
                    const float *bias = s_bias + 64 * l;
#pragma unroll
                    for (int kb = 0; kb < 64; kb += 4) {
                        float4 o;
                        float *op = &o.x;
                        float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (P.use_res && l > 0) res = *reinterpret_cast<const float4 *>(act + canon_off((uint32_t)row, kb, 64));  // MLP:237-239
                        const float *rp = &res.x;
#pragma unroll
                        for (int j = 0; j < 4; ++j) op[j] = tanh_fast(v[kb + j] + bias[kb + j] + rp[j]);
                        *reinterpret_cast<float4 *>(act + canon_off((uint32_t)row, kb, 64)) = o;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
            }
            // ---- output layer: one qudit per chunk, MMA of chunk c+1 under the epilogue of chunk c -------------------------
            if (tid == 0) {
                mbar_wait(&bar_wc[0], p_wc[0]);
                tc_fence_after();
                for (int t = 0; t < live_tiles; ++t) issue_gemm(tmem + 128u * t, smem_u32(acts + t * TC_ACT_BYTES), smem_u32(wc), 64);
                umma_commit(&bar_m[0]);
            }
            for (int c = 0; c < nchunks; ++c) {
                const int b = c & 1;
                if (tid == 0 && c + 1 < nchunks) {
                    mbar_wait(&bar_wc[b ^ 1], p_wc[b ^ 1]);
                    tc_fence_after();
                    for (int t = 0; t < live_tiles; ++t)
                        issue_gemm(tmem + 128u * t + 64u * (b ^ 1), smem_u32(acts + t * TC_ACT_BYTES), smem_u32(wc) + (b ^ 1) * TC_WH_BYTES, 64);
                    umma_commit(&bar_m[b ^ 1]);
                }
                p_wc[b] ^= 1u;
                mbar_wait(&bar_m[b], p_m[b]);
                p_m[b] ^= 1u;
                tc_fence_after();
                if (tl < live_tiles) {
                    const int q = q_lo + c;
                    float z[64];
                    tmem_ld64(my_tmem + 64u * b, z);
                    const float *bias = s_bias + 64 * depth + 64 * q;
                    const int start = P.qudit_starts[q], bits = P.qudit_starts[q + 1] - start;
                    const int chosen = (int)((x >> start) & ((1ull << bits) - 1ull));
                    const int DM = P.max_qudit_dim;
                    if (net == 0) {
                        uint64_t mw;
                        if (P.du[q]) {  // unmasked qudit: max_qudit_dim columns for log psi, the qudit's own 2^bits for the samplers (k3_made.cu)
                            const int dq = MODE == 1 ? (1 << bits) : DM;
                            mw = dq >= 64 ? ~0ull : ((1ull << dq) - 1ull);
                        } else {
                            const uint64_t prefix = start == 0 ? 0ull : (x & ((1ull << start) - 1ull));
                            const long long mi = tc_memo_index(P, prefix);
                            mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)q * P.memo_size + mi) : 0ull;
                        }
                        const uint32_t mlo = (uint32_t)mw, mhi = (uint32_t)(mw >> 32);
                        float sum = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) {
                            z[d] += bias[d];
                            sum += d < DM ? z[d] : 0.f;
                        }
                        const float mean = P.subtract_mean ? sum / (float)DM : 0.f;  // before masking (ANQS:338-340)
                        float mx = -INFINITY;
#pragma unroll
                        for (int d = 0; d < 64; ++d) {
                            const bool ok = ((d < 32 ? mlo >> d : mhi >> (d - 32)) & 1u) != 0u;
                            z[d] = ok ? z[d] - mean : -INFINITY;   // masked logits: exp() of them is exactly 0
                            mx = fmaxf(mx, z[d]);
                        }
                        const bool any = mw != 0ull;
                        const float mxs = any ? mx : 0.f;
                        float se = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d) se += __expf(2.0f * (z[d] - mxs));
                        const float Lnorm = mxs + 0.5f * __logf(se);  // 0.5 * logsumexp(2 z) over the allowed outcomes
                        if (MODE == 1) {
                            if (s < B) {
#pragma unroll
                                for (int d = 0; d < 64; ++d)
                                    if (d < DM) cond_out[(size_t)s * DM + d] = any ? (double)(z[d] - Lnorm) : -INFINITY;
                            }
                        } else {
                            float pick = 0.f;
#pragma unroll
                            for (int d = 0; d < 64; ++d)
                                if (d == chosen) pick = z[d];
                            if (any && pick > -INFINITY)
                                acc_re += (double)(pick - Lnorm);
                            else
                                dead = true;
                        }
                    } else {
                        float pick = 0.f;
#pragma unroll
                        for (int d = 0; d < 64; ++d)
                            if (d == chosen) pick = z[d] + bias[d];
                        acc_im += (double)pick;
                    }
                }
                tc_fence_before();
                __syncthreads();  // everyone has read TMEM buffer b; its MMA is done, so weight buffer b is free
                if (tid == 0 && c + 2 < nchunks) {
                    mbar_arrive_expect_tx(&bar_wc[b], TC_WH_BYTES);
                    bulk_copy_g2s(wc + b * TC_WH_BYTES, packed + L.w3[net] + (uint32_t)(q_lo + c + 2) * TC_WH_BYTES, TC_WH_BYTES, &bar_wc[b]);
                }
            }
        }
        if (MODE == 0 && s < B)
            log_psi[s] = dead ? make_double2(-INFINITY, 0.0) : make_double2(acc_re, 3.14159265358979323846 * acc_im);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

}  // namespace anqs

using namespace anqs;

static int tc_check(const anqs_made_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->qudit_num >= 1 && P->qudit_num <= 64, "qudit_num must be in [1, 64]");
    ANQS_REQUIRE(P->max_qudit_dim >= 2 && P->max_qudit_dim <= 64, "max_qudit_dim must be in [2, 64]");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] hidden layers");
    ANQS_REQUIRE(P->width == TC_W, "hidden width must be 64 (the reference default)");
    for (int l = 0; l <= P->depth; ++l) ANQS_REQUIRE(P->w_abs[l] && P->w_phase[l], "null weight pointer");
    ANQS_REQUIRE(P->cont_mask && P->memo_size >= 1, "null continuation-mask table");
    return 0;
}

static size_t tc_smem_bytes(const anqs_made_desc_t *P) {
    return TC_NT * (size_t)TC_ACT_BYTES + 3 * (size_t)TC_WH_BYTES + (size_t)(64 * P->depth + P->qudit_num * 64) * 4;
}

extern "C" {

size_t anqs_made_tc_packed_bytes(const anqs_made_desc_t *desc) {
    if (!desc) return 0;
    return tc_layout(desc->qubit_num, desc->qudit_num, desc->depth).total;
}

int anqs_made_tc_pack(const anqs_made_desc_t *desc, void *d_packed, void *stream) {
    if (tc_check(desc)) return 1;
    ANQS_REQUIRE(d_packed && ((uintptr_t)d_packed & 127) == 0, "packed buffer must be non-null and 128-byte aligned");
    TcPacked L = tc_layout(desc->qubit_num, desc->qudit_num, desc->depth);
    made_tc_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*desc, L, (unsigned char *)d_packed);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_made_log_psi_tc(const anqs_made_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n,
                         double *d_log_psi, void *stream) {
    if (tc_check(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_idx && d_log_psi, "null pointer");
    TcPacked L = tc_layout(desc->qubit_num, desc->qudit_num, desc->depth);
    const size_t smem = tc_smem_bytes(desc);
    auto kern = made_tc_kernel<0>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ngroups = (n + 128 * TC_NT - 1) / (128 * TC_NT);
    const int grid = (int)std::min<int64_t>(ngroups, sm_count_of_current_device());
    kern<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(*desc, L, (const unsigned char *)d_packed, d_idx, n, 0,
                                                          (double2 *)d_log_psi, nullptr);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_made_cond_log_abs_tc(const anqs_made_desc_t *desc, const void *d_packed, int qudit_idx, const int64_t *d_prefix,
                              int64_t n, double *d_cond, void *stream) {
    if (tc_check(desc)) return 1;
    ANQS_REQUIRE(qudit_idx >= 0 && qudit_idx < desc->qudit_num, "qudit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_prefix && d_cond, "null pointer");
    TcPacked L = tc_layout(desc->qubit_num, desc->qudit_num, desc->depth);
    const size_t smem = tc_smem_bytes(desc);
    auto kern = made_tc_kernel<1>;
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ngroups = (n + 128 * TC_NT - 1) / (128 * TC_NT);
    const int grid = (int)std::min<int64_t>(ngroups, sm_count_of_current_device());
    kern<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(*desc, L, (const unsigned char *)d_packed, d_prefix, n, qudit_idx,
                                                          nullptr, d_cond);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
