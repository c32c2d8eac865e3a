// Kernel family 5, tensor-core mode: the autoregressive transformer wave function (k5_transformer.cu, reference
// legacy/anqs_primitives/made/transformer_made.py:9-48) with every projection on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, fp32 accumulators in TMEM) and everything between the GEMMs - bias, causal multi-head attention,
// residual + LayerNorm, ReLU, decoder, symmetry masks, normalisation, gather - in fp32 registers of the thread that owns the
// token.  Inference only; agreement with the fp64 kernel is a stated tolerance (tests/test_gpu_transformer.py), not 1e-10.
//
// One CTA = 128 threads = one tile of 128 token rows = floor(128 / T) samples of T tokens (BOS + known bits); thread r owns
// row r: TMEM lane r, 64 residual-stream values in registers.
//   per layer: x -> A operand (canonical K-major layout in shared memory) -> 3 MMA groups (Q, K, V; 192 TMEM columns)
//              -> Q into registers, K and V rows into shared memory -> online-softmax attention over the <= T causal keys of
//              the thread's own sample -> A -> MMA (out_proj) -> x = LN1(x + .) -> A -> MMA (linear1) -> ReLU -> A -> MMA
//              (linear2) -> x = LN2(x + .)
//   weights:   packed once per parameter update (transformer_tc_pack_kernel): the six 64 x 64 matrices of a layer in the
//              MMA's operand layout + its fp32 vectors, staged per layer by bulk TMA copies (96 KB + 2.5 KB)
//   epilogue:  decoder (4 dot products per token), continuation mask of the token's prefix, 0.5 logsumexp(2 re) over the
//              allowed outcomes; the T contributions of a sample are summed through shared memory.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace anqs {

constexpr int TFC_THREADS = 128;
constexpr int TFC_D = 64;
constexpr uint32_t TFC_MAT = 64 * 64 * 4;             // one packed weight matrix
constexpr uint32_t TFC_VEC_FLOATS = 640;              // bq bk bv | bo | b1 | b2 | ln1w ln1b ln2w ln2b
constexpr uint32_t TFC_LAYER_BYTES = 6 * TFC_MAT + TFC_VEC_FLOATS * 4;
constexpr int TFC_KV_STRIDE = 68;                     // floats per staged K / V row: 16-byte aligned rows for float4 reads, 4-bank skew between rows

struct TfcLayout {
    uint32_t layer[4];     // per layer: [Wq Wk Wv Wo W1 W2][vectors]
    uint32_t tok, pos, decw, decb, total;
};

__host__ __device__ inline TfcLayout tfc_layout(int qubit_num, int depth) {
    TfcLayout L;
    uint32_t off = 0;
    for (int l = 0; l < 4; ++l) {
        L.layer[l] = off;
        if (l < depth) off += TFC_LAYER_BYTES;
    }
    L.tok = off; off += 3 * TFC_D * 4;
    L.pos = off; off += (uint32_t)(qubit_num + 1) * TFC_D * 4;
    L.decw = off; off += 4 * TFC_D * 4;
    L.decb = off; off += 16;
    L.total = (off + 127) / 128 * 128;
    return L;
}

__global__ void transformer_tc_pack_kernel(const anqs_transformer_desc_t P, TfcLayout L, unsigned char *out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int l = 0; l < P.depth; ++l) {
        unsigned char *base = out + L.layer[l];
        for (int64_t e = t0; e < 6 * 64 * 64; e += stride) {
            const int m = (int)(e >> 12), j = (int)((e >> 6) & 63), k = (int)(e & 63);
            const double *W = m < 3 ? P.in_proj_w[l] + (size_t)m * 64 * 64 : m == 3 ? P.out_proj_w[l] : m == 4 ? P.lin1_w[l] : P.lin2_w[l];
            *reinterpret_cast<float *>(base + (uint32_t)m * TFC_MAT + canon_off((uint32_t)j, (uint32_t)k, 64)) = (float)W[(size_t)j * 64 + k];
        }
        float *vec = reinterpret_cast<float *>(base + 6 * TFC_MAT);
        for (int64_t e = t0; e < TFC_VEC_FLOATS; e += stride) {
            const int i = (int)e;
            double v;
            if (i < 192) v = P.in_proj_b[l] ? P.in_proj_b[l][i] : 0.0;
            else if (i < 256) v = P.out_proj_b[l] ? P.out_proj_b[l][i - 192] : 0.0;
            else if (i < 320) v = P.lin1_b[l] ? P.lin1_b[l][i - 256] : 0.0;
            else if (i < 384) v = P.lin2_b[l] ? P.lin2_b[l][i - 320] : 0.0;
            else if (i < 448) v = P.ln1_w[l][i - 384];
            else if (i < 512) v = P.ln1_b[l][i - 448];
            else if (i < 576) v = P.ln2_w[l][i - 512];
            else v = P.ln2_b[l][i - 576];
            vec[i] = (float)v;
        }
    }
    for (int64_t e = t0; e < 3 * TFC_D; e += stride) reinterpret_cast<float *>(out + L.tok)[e] = (float)P.tok_emb[e];
    for (int64_t e = t0; e < (int64_t)(P.qubit_num + 1) * TFC_D; e += stride) reinterpret_cast<float *>(out + L.pos)[e] = (float)P.pos_emb[e];
    for (int64_t e = t0; e < 4 * TFC_D; e += stride) reinterpret_cast<float *>(out + L.decw)[e] = (float)P.dec_w[e];
    for (int64_t e = t0; e < 4; e += stride) reinterpret_cast<float *>(out + L.decb)[e] = (float)P.dec_b[e];
}

// x -> the thread's row of the A operand
__device__ __forceinline__ void tfc_store_row(unsigned char *A, int row, const float (&x)[64]) {
#pragma unroll
    for (int kb = 0; kb < 64; kb += 4)
        *reinterpret_cast<float4 *>(A + canon_off((uint32_t)row, (uint32_t)kb, 64)) = make_float4(x[kb], x[kb + 1], x[kb + 2], x[kb + 3]);
}

// x <- LayerNorm(x + y + bias) over the 64 columns (biased variance, eps inside the square root)
__device__ __forceinline__ void tfc_residual_ln(float (&x)[64], const float (&y)[64], const float *bias, const float *gamma,
                                                const float *beta, float eps) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        x[k] += y[k] + bias[k];
        sum += x[k];
    }
    const float mean = sum * (1.0f / 64.0f);
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        x[k] -= mean;
        sq += x[k] * x[k];
    }
    const float rstd = rsqrtf(sq * (1.0f / 64.0f) + eps);
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = x[k] * rstd * gamma[k] + beta[k];
}

// MODE 0: log psi of whole configurations.  MODE 1: normalised conditional log|psi| of qubit `level` for prefixes.
template <int MODE, int HD>
__global__ void __launch_bounds__(TFC_THREADS, 1)
transformer_tc_kernel(const anqs_transformer_desc_t P, const TfcLayout L, const unsigned char *__restrict__ packed,
                      const int64_t *__restrict__ idx_in, int64_t B, int level, double2 *__restrict__ log_psi,
                      double *__restrict__ cond_out) {
    extern __shared__ __align__(1024) unsigned char tfc_smem[];
    unsigned char *A = tfc_smem;                                            // 32 KB activation operand
    unsigned char *W = A + 128 * 64 * 4;                                    // one layer: six matrices + vectors
    float *vec = reinterpret_cast<float *>(W + 6 * TFC_MAT);
    float *Ks = reinterpret_cast<float *>(W + TFC_LAYER_BYTES);             // [128][65]
    float *Vs = Ks + 128 * TFC_KV_STRIDE;
    float *s_dec = Vs + 128 * TFC_KV_STRIDE;                                // decoder weights [4][64] + bias [4]
    float *s_part = s_dec + 4 * TFC_D + 4;                                  // per row: (re, im, dead)
    __shared__ uint64_t bar_w, bar_m;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, row = tid;
    const int n = P.qubit_num, depth = P.depth;
    const int T = MODE == 1 ? level + 1 : n;   // tokens per sample whose outputs are needed (BOS + known bits)
    const int S = 128 / T;                      // samples per tile
    const float scale = rsqrtf((float)HD);
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_m, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < 4 * TFC_D + 4; e += TFC_THREADS)
        s_dec[e] = e < 4 * TFC_D ? reinterpret_cast<const float *>(packed + L.decw)[e] : reinterpret_cast<const float *>(packed + L.decb)[e - 4 * TFC_D];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);  // this thread's lane, column 0
    uint32_t p_w = 0, p_m = 0;
    const float *tok_emb = reinterpret_cast<const float *>(packed + L.tok), *pos_emb = reinterpret_cast<const float *>(packed + L.pos);

    const int64_t ntiles = (B + S - 1) / S;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s_loc = row / T, t = row - s_loc * T;
        const int64_t smp = tile * S + s_loc;
        const bool live = s_loc < S && smp < B;
        const uint64_t xbits = live ? (uint64_t)idx_in[smp] : 0ull;
        const int r0 = s_loc * T;  // first row of this thread's sample
        // ---- embedding: token (BOS = 2, then the bits) + position -------------------------------------------------------------
        float x[64];
        {
            const int tok = t == 0 ? 2 : (int)((xbits >> (t - 1)) & 1ull);
            const float4 *te = reinterpret_cast<const float4 *>(tok_emb + tok * TFC_D), *pe = reinterpret_cast<const float4 *>(pos_emb + t * TFC_D);
#pragma unroll
            for (int k4 = 0; k4 < 16; ++k4) {
                const float4 a = __ldg(te + k4), b = __ldg(pe + k4);
                x[4 * k4] = live ? a.x + b.x : 0.f;
                x[4 * k4 + 1] = live ? a.y + b.y : 0.f;
                x[4 * k4 + 2] = live ? a.z + b.z : 0.f;
                x[4 * k4 + 3] = live ? a.w + b.w : 0.f;
            }
        }
        for (int l = 0; l < depth; ++l) {
            __syncthreads();  // the previous users of A / W / Ks / Vs are done
            if (tid == 0) {
                mbar_arrive_expect_tx(&bar_w, TFC_LAYER_BYTES);
                for (uint32_t off = 0; off < TFC_LAYER_BYTES; off += 32768u)
                    bulk_copy_g2s(W + off, packed + L.layer[l] + off, min(32768u, TFC_LAYER_BYTES - off), &bar_w);
            }
            tfc_store_row(A, row, x);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- q, k, v projections: TMEM columns [0,64), [64,128), [128,192) -----------------------------------------------
            if (tid == 0) {
                mbar_wait(&bar_w, p_w);
                tc_fence_after();
                for (uint32_t m = 0; m < 3; ++m) issue_gemm(tmem + 64u * m, smem_u32(A), smem_u32(W) + m * TFC_MAT, 64);
                umma_commit(&bar_m);
            }
            mbar_wait(&bar_w, p_w);  // everyone reads the layer's vectors below
            p_w ^= 1u;
            mbar_wait(&bar_m, p_m);
            p_m ^= 1u;
            tc_fence_after();
            float q[64];
            {
                float kv[64];
                tmem_ld64(my_tmem + 64u, kv);
#pragma unroll
                for (int k = 0; k < 64; ++k) Ks[row * TFC_KV_STRIDE + k] = kv[k] + vec[64 + k];
                tmem_ld64(my_tmem + 128u, kv);
#pragma unroll
                for (int k = 0; k < 64; ++k) Vs[row * TFC_KV_STRIDE + k] = kv[k] + vec[128 + k];
                tmem_ld64(my_tmem, q);
#pragma unroll
                for (int k = 0; k < 64; ++k) q[k] = (q[k] + vec[k]) * scale;
            }
            tc_fence_before();
            __syncthreads();
            // ---- causal attention over the thread's own sample: online softmax per head, result overwrites q ----------------
            // (HD and the head loop are compile-time so that q[] stays in registers)
#pragma unroll
            for (int h = 0; h < TFC_D / HD; ++h) {
                float m_run = -INFINITY, den = 0.f, o[HD];
#pragma unroll
                for (int d = 0; d < HD; ++d) o[d] = 0.f;
                for (int tp = 0; tp <= t; ++tp) {
                    const float4 *kr = reinterpret_cast<const float4 *>(Ks + (r0 + tp) * TFC_KV_STRIDE + h * HD);
                    const float4 *vr = reinterpret_cast<const float4 *>(Vs + (r0 + tp) * TFC_KV_STRIDE + h * HD);
                    float sc = 0.f;
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 kk = kr[d4];
                        sc = fmaf(q[h * HD + 4 * d4], kk.x, sc);
                        sc = fmaf(q[h * HD + 4 * d4 + 1], kk.y, sc);
                        sc = fmaf(q[h * HD + 4 * d4 + 2], kk.z, sc);
                        sc = fmaf(q[h * HD + 4 * d4 + 3], kk.w, sc);
                    }
                    const float m_new = fmaxf(m_run, sc);
                    const float corr = __expf(m_run - m_new), pw = __expf(sc - m_new);
                    den = den * corr + pw;
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 vv = vr[d4];
                        o[4 * d4] = o[4 * d4] * corr + pw * vv.x;
                        o[4 * d4 + 1] = o[4 * d4 + 1] * corr + pw * vv.y;
                        o[4 * d4 + 2] = o[4 * d4 + 2] * corr + pw * vv.z;
                        o[4 * d4 + 3] = o[4 * d4 + 3] * corr + pw * vv.w;
                    }
                    m_run = m_new;
                }
                const float inv = 1.0f / den;
#pragma unroll
                for (int d = 0; d < HD; ++d) q[h * HD + d] = o[d] * inv;
            }
            // ---- output projection + residual + LayerNorm 1 ---------------------------------------------------------------------
            __syncthreads();  // every thread is done reading Ks / Vs and its own A row was consumed by the q, k, v MMAs
            tfc_store_row(A, row, q);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_gemm(tmem, smem_u32(A), smem_u32(W) + 3 * TFC_MAT, 64);
                umma_commit(&bar_m);
            }
            mbar_wait(&bar_m, p_m);
            p_m ^= 1u;
            tc_fence_after();
            tmem_ld64(my_tmem, q);
            tfc_residual_ln(x, q, vec + 192, vec + 384, vec + 448, (float)P.ln_eps);
            // ---- feed-forward (dim -> dim, ReLU, dim -> dim) + residual + LayerNorm 2 ----------------------------------------------
            tc_fence_before();
            __syncthreads();  // all TMEM reads of the previous accumulator are done before the next MMA overwrites it
            tfc_store_row(A, row, x);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_gemm(tmem, smem_u32(A), smem_u32(W) + 4 * TFC_MAT, 64);
                umma_commit(&bar_m);
            }
            mbar_wait(&bar_m, p_m);
            p_m ^= 1u;
            tc_fence_after();
            tmem_ld64(my_tmem, q);
#pragma unroll
            for (int k = 0; k < 64; ++k) q[k] = fmaxf(q[k] + vec[256 + k], 0.f);
            tc_fence_before();
            __syncthreads();
            tfc_store_row(A, row, q);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_gemm(tmem, smem_u32(A), smem_u32(W) + 5 * TFC_MAT, 64);
                umma_commit(&bar_m);
            }
            mbar_wait(&bar_m, p_m);
            p_m ^= 1u;
            tc_fence_after();
            tmem_ld64(my_tmem, q);
            tfc_residual_ln(x, q, vec + 320, vec + 512, vec + 576, (float)P.ln_eps);
            tc_fence_before();
        }
        // ---- decoder, masks, normalisation: the token's own contribution ------------------------------------------------------
        float o4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float acc = s_dec[4 * TFC_D + c];
#pragma unroll
            for (int k = 0; k < 64; ++k) acc = fmaf(x[k], s_dec[c * TFC_D + k], acc);
            o4[c] = acc;
        }
        float re = 0.f, im = 0.f, dead = 0.f;
        if (live && (MODE == 0 || t == level)) {
            const uint64_t prefix = t == 0 ? 0ull : (xbits & ((1ull << t) - 1ull));
            long long mi = 0;
            for (int sy = 0; sy < P.sym_num; ++sy) {
                const int64_t *d = P.sym[sy];
                long long ev;
                if (d[0] == 0) ev = d[7] + __popcll(prefix & (uint64_t)d[1]) - __popcll(prefix & (uint64_t)d[2]);
                else ev = (__popcll(prefix & (uint64_t)d[1]) & 1) ? -d[7] : d[7];
                const long long num = ev * d[3] + d[4], qd = num / d[5], rm = num % d[5];
                mi += ((rm != 0 && ((rm < 0) != (d[5] < 0))) ? qd - 1 : qd) * d[6];
            }
            const uint64_t mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)t * P.memo_size + mi) : 0ull;
            const bool a0 = mw & 1ull, a1 = (mw >> 1) & 1ull;
            const float z0 = a0 ? o4[0] : -INFINITY, z1 = a1 ? o4[2] : -INFINITY;
            const float mx = fmaxf(z0, z1);
            const float Ln = mx + 0.5f * __logf((a0 ? __expf(2.0f * (z0 - mx)) : 0.f) + (a1 ? __expf(2.0f * (z1 - mx)) : 0.f));
            if (MODE == 1) {
                cond_out[(size_t)smp * 2 + 0] = a0 ? (double)(z0 - Ln) : -INFINITY;
                cond_out[(size_t)smp * 2 + 1] = a1 ? (double)(z1 - Ln) : -INFINITY;
            } else {
                const int bit = (int)((xbits >> t) & 1ull);
                if (bit ? a1 : a0) {
                    re = (bit ? z1 : z0) - Ln;
                    im = bit ? o4[3] : o4[1];
                } else {
                    dead = 1.f;
                }
            }
        }
        if (MODE == 0) {
            __syncthreads();  // s_part of the previous tile has been consumed
            s_part[row * 3] = re;
            s_part[row * 3 + 1] = im;
            s_part[row * 3 + 2] = dead;
            __syncthreads();
            if (live && t == 0) {
                double sr = 0.0, si = 0.0;
                bool dd = false;
                for (int tp = 0; tp < T; ++tp) {
                    sr += (double)s_part[(r0 + tp) * 3];
                    si += (double)s_part[(r0 + tp) * 3 + 1];
                    dd = dd || s_part[(r0 + tp) * 3 + 2] != 0.f;
                }
                log_psi[smp] = dd ? make_double2(-INFINITY, 0.0) : make_double2(sr, si);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static size_t tfc_smem_bytes() {
    return (size_t)128 * 64 * 4 + TFC_LAYER_BYTES + (size_t)2 * 128 * TFC_KV_STRIDE * 4 + (4 * TFC_D + 4) * 4 + 128 * 3 * 4 + 64;
}

}  // namespace anqs

using namespace anqs;

static int tfc_check(const anqs_transformer_desc_t *P) {
    ANQS_REQUIRE(P, "null network descriptor");
    ANQS_REQUIRE(P->qubit_num >= 1 && P->qubit_num <= 64, "qubit_num must be in [1, 64]");
    ANQS_REQUIRE(P->dim == TFC_D, "model dimension must be 64");
    ANQS_REQUIRE(P->depth >= 1 && P->depth <= 4, "depth must be in [1, 4] encoder layers");
    ANQS_REQUIRE(P->head_num == 4 || P->head_num == 8 || P->head_num == 16,
                 "the tensor-core mode keeps a head in registers: head_num must be 4, 8 or 16 (head dimension <= 16)");
    ANQS_REQUIRE(P->tok_emb && P->pos_emb && P->dec_w && P->dec_b, "null embedding / decoder pointer");
    for (int l = 0; l < P->depth; ++l)
        ANQS_REQUIRE(P->in_proj_w[l] && P->out_proj_w[l] && P->lin1_w[l] && P->lin2_w[l] && P->ln1_w[l] && P->ln1_b[l] && P->ln2_w[l] &&
                         P->ln2_b[l], "null layer weight pointer");
    ANQS_REQUIRE(P->cont_mask && P->memo_size >= 1, "null continuation-mask table");
    return 0;
}

template <int MODE>
static int tfc_launch(const anqs_transformer_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n, int level,
                      double *d_log_psi, double *d_cond, void *stream) {
    auto kern = desc->head_num == 4 ? transformer_tc_kernel<MODE, 16> : desc->head_num == 8 ? transformer_tc_kernel<MODE, 8>
                                                                                            : transformer_tc_kernel<MODE, 4>;
    const size_t smem = tfc_smem_bytes();
    ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int T = MODE == 1 ? level + 1 : desc->qubit_num;
    const int S = 128 / T;
    const int64_t ntiles = (n + S - 1) / S;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count_of_current_device());
    kern<<<grid, TFC_THREADS, smem, (cudaStream_t)stream>>>(*desc, tfc_layout(desc->qubit_num, desc->depth), (const unsigned char *)d_packed,
                                                           d_idx, n, level, (double2 *)d_log_psi, d_cond);
    ANQS_LAUNCH_CHECK();
    return 0;
}

extern "C" {

size_t anqs_transformer_tc_packed_bytes(const anqs_transformer_desc_t *desc) {
    if (!desc) return 0;
    return tfc_layout(desc->qubit_num, desc->depth).total;
}

int anqs_transformer_tc_pack(const anqs_transformer_desc_t *desc, void *d_packed, void *stream) {
    if (tfc_check(desc)) return 1;
    ANQS_REQUIRE(d_packed && ((uintptr_t)d_packed & 127) == 0, "packed buffer must be non-null and 128-byte aligned");
    transformer_tc_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(*desc, tfc_layout(desc->qubit_num, desc->depth), (unsigned char *)d_packed);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_transformer_log_psi_tc(const anqs_transformer_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n,
                                double *d_log_psi, void *stream) {
    if (tfc_check(desc)) return 1;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_idx && d_log_psi, "null pointer");
    return tfc_launch<0>(desc, d_packed, d_idx, n, 0, d_log_psi, nullptr, stream);
}

int anqs_transformer_cond_log_abs_tc(const anqs_transformer_desc_t *desc, const void *d_packed, int qubit_idx, const int64_t *d_prefix,
                                     int64_t n, double *d_cond, void *stream) {
    if (tfc_check(desc)) return 1;
    ANQS_REQUIRE(qubit_idx >= 0 && qubit_idx < desc->qubit_num, "qubit index out of range");
    ANQS_REQUIRE(n >= 0, "negative prefix count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_packed && d_prefix && d_cond, "null pointer");
    return tfc_launch<1>(desc, d_packed, d_prefix, n, qubit_idx, nullptr, d_cond, stream);
}

}  // extern "C"
