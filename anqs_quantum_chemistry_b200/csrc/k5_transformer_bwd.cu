// Kernel family 5, backward: the per-sample gradient chain of the transformer wave function (k5_transformer.cu), float64.
// Reference: autograd through TransformerMADE (legacy/anqs_primitives/made/transformer_made.py:9-48 = nn.TransformerEncoder of
// post-norm layers) and the masked normalisation ANQS:392-405.  Same split as the MADE backward (k3_made_bwd.cu): this kernel
// does everything that is per sample - it recomputes the forward pass of a tile of floor(64 / T) samples, keeps what the
// backward needs in a per-row workspace, and walks the layers back (decoder, LayerNorm, feed-forward, LayerNorm, attention,
// embeddings) writing, per row, the gradient signal of every linear layer next to that layer's input.  The reductions over
// the batch - grad W = dY^T X, grad b = column sums - are then the ordinary problems of batch_reduce_gemm_kernel
// (k3_batch_reduce.cu); the embedding gradients are one more such problem against a one-hot (token | position) matrix; the
// LayerNorm weight / bias gradients are column sums accumulated per CTA in a fixed order and added up by a small second kernel.
//
// Layout as in the forward kernel: activations [64 columns][rows] in shared memory (stride MD_S), a thread owns rows
// ty*4..ty*4+3 and columns tx, tx+16, tx+32, tx+48; every projection and its transpose is one 64 x 64 x 64 DFMA tile.
#include <algorithm>

#include "common.cuh"
#include "made_common.cuh"

namespace anqs {

constexpr int TB_D = 64;
constexpr int TB_MAXH = 16;                                  // heads: 1, 2, 4, 8 or 16 (the head dimension is a template parameter)
constexpr int TB_BUF = 64 * MD_S;                            // doubles per activation buffer
constexpr int TB_NVEC = 16;                                  // LayerNorm vectors: 4 per layer (g1, b1, g2, b2), depth <= 4
constexpr size_t TB_SMEM = (size_t)(5 * TB_BUF + 3 * 64 * TB_MAXH + 16 * 64 + TB_NVEC * 64 + 64 * 4) * sizeof(double) + 64 * sizeof(uint64_t);

struct TfBwdPtrs {   // per-row workspace of one chunk of samples (R = samples * T rows), all [R][64] unless said otherwise
    double *xin[4], *q[4], *k[4], *v[4], *a[4], *y1[4], *x1[4], *hf[4], *y2[4], *xf;
    double *gqkv[4];   // [R][192]
    double *gy1[4], *ghp[4], *gy2[4];
    double *gdec;      // [R][4]
    double *gx0;       // [R][64]
    double *emb;       // [R][P] one-hot: column tok (0..2) and column 3 + t
    double *vec;       // [grid][TB_NVEC][64] per-CTA LayerNorm sums
    int P;
};

// Tile <-> workspace rows.  A warp moves blocks of 4 rows x 16 columns, a lane two consecutive rows of one column
// (lane = column * 2 + row pair): one 16-byte shared-memory access - the 8 lanes of a quarter-warp touch 8 different
// 16-byte bank groups, (2 * column + row pair) mod 8 with the column stride MD_S = 68 doubles - and two 8-byte global
// accesses that, over the warp, cover two full 128-byte row segments each.
// dst[j][r] = src[(row0 + r) * 64 + j], zero beyond `rows`
__device__ __forceinline__ void load_rows(double *dst, const double *__restrict__ src, int64_t row0, int rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rp = lane & 1, jl = lane >> 1;
    for (int b = warp; b < 64; b += MD_THREADS / 32) {
        const int r = (b >> 2) * 4 + rp * 2, j = (b & 3) * 16 + jl;
        const double *g = src + (row0 + r) * 64 + j;
        double2 v;
        v.x = r < rows ? g[0] : 0.0;
        v.y = r + 1 < rows ? g[64] : 0.0;
        *reinterpret_cast<double2 *>(dst + j * MD_S + r) = v;
    }
}
__device__ __forceinline__ void store_rows(const double *src, double *__restrict__ dst, int64_t row0, int rows, int ld = 64, int col0 = 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, rp = lane & 1, jl = lane >> 1;
    for (int b = warp; b < 64; b += MD_THREADS / 32) {
        const int r = (b >> 2) * 4 + rp * 2, j = (b & 3) * 16 + jl;
        const double2 v = *reinterpret_cast<const double2 *>(src + j * MD_S + r);
        double *g = dst + (row0 + r) * ld + col0 + j;
        if (r < rows) g[0] = v.x;
        if (r + 1 < rows) g[ld] = v.y;
    }
}
// Weight tiles travel global -> registers -> shared memory in two steps, so that the loads of the NEXT tile are in flight
// while the current one is multiplied (one CTA per SM: nothing else would hide their latency).
//   transposed (forward):  wt[k][j] = W[(row0 + j) * 64 + k]   (nn.Linear layout [out][in]; mapping of load_weights_t)
//   plain (backward):      wt[j][k] = W[(row0 + j) * 64 + k]   (out[r][k] = sum_j in[r][j] W[row0 + j][k]: multiplication by W)
__device__ __forceinline__ void fetch_w_t(double (&v)[16], const double *__restrict__ W, int row0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kq = lane & 3, jo = lane >> 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        v[i] = __ldg(W + (size_t)(row0 + j) * 64 + k);
    }
}
__device__ __forceinline__ void commit_w_t(double *wt, const double (&v)[16]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kq = lane & 3, jo = lane >> 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int t = warp + (MD_THREADS / 32) * i, j = (t & 7) * 8 + jo, k = (t >> 3) * 4 + kq;
        wt[k * MD_S + j] = v[i];
    }
}
__device__ __forceinline__ void fetch_w_n(double (&v)[16], const double *__restrict__ W, int row0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = threadIdx.x + MD_THREADS * i;
        v[i] = __ldg(W + (size_t)(row0 + (e >> 6)) * 64 + (e & 63));
    }
}
__device__ __forceinline__ void commit_w_n(double *wt, const double (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int e = threadIdx.x + MD_THREADS * i;
        wt[(e >> 6) * MD_S + (e & 63)] = v[i];
    }
}
// vec[j] += sum over the tile's rows of c (each thread brings the sums over its own 4 rows for its 4 columns); fixed order
__device__ __forceinline__ void col_reduce_add(double *vec, double *red, const double (&c)[4], int tx, int ty) {
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) red[ty * 64 + tx + 16 * jj] = c[jj];
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
        for (int y = 0; y < 16; ++y) s += red[y * 64 + threadIdx.x];
        vec[threadIdx.x] += s;
    }
}

// LayerNorm backward in place: G holds dOut on entry and d(pre-norm value) on exit; Y = the pre-norm values of the tile.
// Adds the tile's column sums of dOut * xhat and dOut to vec_g / vec_b.  Rows beyond `rows` hold zeros and stay zero.
__device__ __forceinline__ void layer_norm_backward(double *G, const double *Y, const double *__restrict__ gamma, double eps, double *vec_g,
                                                    double *vec_b, double *red, int rows, int tx, int ty) {
    double cg[4] = {0.0, 0.0, 0.0, 0.0}, cb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) {
        const int r = ty * 4 + ss;
        double v[4], sum = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            v[jj] = Y[(tx + 16 * jj) * MD_S + r];
            sum += v[jj];
        }
        const double mean = row_sum16(sum) * (1.0 / 64.0);
        double sq = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            v[jj] -= mean;
            sq += v[jj] * v[jj];
        }
        const double rstd = 1.0 / sqrt(row_sum16(sq) * (1.0 / 64.0) + eps);
        double dg[4], m1 = 0.0, m2 = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tx + 16 * jj;
            const double xh = v[jj] * rstd, d = G[j * MD_S + r];
            v[jj] = xh;
            if (r < rows) {
                cg[jj] += d * xh;
                cb[jj] += d;
            }
            dg[jj] = d * __ldg(gamma + j);
            m1 += dg[jj];
            m2 += dg[jj] * xh;
        }
        m1 = row_sum16(m1) * (1.0 / 64.0);
        m2 = row_sum16(m2) * (1.0 / 64.0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) G[(tx + 16 * jj) * MD_S + r] = r < rows ? rstd * (dg[jj] - m1 - v[jj] * m2) : 0.0;
    }
    col_reduce_add(vec_g, red, cg, tx, ty);
    col_reduce_add(vec_b, red, cb, tx, ty);
    __syncthreads();
}

// x <- LayerNorm(x + acc + bias); the pre-norm value goes to `pre` ([col][row] shared buffer)
__device__ __forceinline__ void residual_layer_norm_save(double *x, double *pre, const double (&acc)[4][4], const double *__restrict__ bias,
                                                         const double *__restrict__ gamma, const double *__restrict__ beta, double eps,
                                                         int tx, int ty) {
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) {
        const int r = ty * 4 + ss;
        double v[4], sum = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tx + 16 * jj;
            v[jj] = acc[ss][jj] + (bias ? __ldg(bias + j) : 0.0) + x[j * MD_S + r];
            pre[j * MD_S + r] = v[jj];
            sum += v[jj];
        }
        const double mean = row_sum16(sum) * (1.0 / 64.0);
        double sq = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            v[jj] -= mean;
            sq += v[jj] * v[jj];
        }
        const double rstd = 1.0 / sqrt(row_sum16(sq) * (1.0 / 64.0) + eps);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = tx + 16 * jj;
            x[j * MD_S + r] = v[jj] * rstd * __ldg(gamma + j) + __ldg(beta + j);
        }
    }
}

__device__ __forceinline__ void store_acc(double *out, const double (&acc)[4][4], const double *__restrict__ bias, int tx, int ty) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = tx + 16 * jj;
        const double b = bias ? __ldg(bias + j) : 0.0;
        double2 *o2 = reinterpret_cast<double2 *>(out + j * MD_S + ty * 4);
        o2[0] = make_double2(acc[0][jj] + b, acc[1][jj] + b);
        o2[1] = make_double2(acc[2][jj] + b, acc[3][jj] + b);
    }
}

// PHASE 0: forward pass that keeps the activations in the workspace and writes log psi (what autograd's forward runs);
// PHASE 1: backward pass from a workspace PHASE 0 filled for the same samples; PHASE 2: both in one launch per chunk of
// samples (the forward pass recomputed - for batches whose activations exceed the caller's workspace).
template <int HD, int PHASE>
__global__ void __launch_bounds__(MD_THREADS, 1)
transformer_backward_kernel(const anqs_transformer_desc_t P, const int64_t *__restrict__ idx_in, int64_t B, const double2 *__restrict__ grad_out,
                            double2 *__restrict__ log_psi, TfBwdPtrs ws) {
    extern __shared__ __align__(16) unsigned char tb_smem[];
    double *buf[5];
    buf[0] = reinterpret_cast<double *>(tb_smem);
    for (int i = 1; i < 5; ++i) buf[i] = buf[i - 1] + TB_BUF;
    double *st_mx = buf[4] + TB_BUF;              // attention statistics per (row, head)
    double *st_den = st_mx + 64 * TB_MAXH;
    double *st_d = st_den + 64 * TB_MAXH;
    double *red = st_d + 64 * TB_MAXH;            // [16][64] scratch of col_reduce_add
    double *vecs = red + 16 * 64;                 // [TB_NVEC][64] LayerNorm sums of this CTA
    double *s_dec = vecs + TB_NVEC * 64;          // decoder outputs, then their gradients [rows][4]
    uint64_t *s_idx = reinterpret_cast<uint64_t *>(s_dec + 64 * 4);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = P.qubit_num, L = P.depth;
    constexpr int H = TB_D / HD, hd = HD;
    const int T = n, S = 64 / T, rows = S * T;
    const double scale = 1.0 / sqrt((double)hd);
    const int64_t ntiles = (B + S - 1) / S;
    for (int e = tid; e < TB_NVEC * 64; e += MD_THREADS) vecs[e] = 0.0;
    double wv[16];           // the weight tile in flight
    bool have_w = false;     // wv holds layer 0's query projection (fetched at the end of the previous tile)

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * S, row0 = base * T;
        const int live = (int)min((int64_t)S, B - base) * T;   // rows that belong to real samples
        double *X = buf[0], *Qb = buf[1], *Kb = buf[2], *Vb = buf[3], *wt = buf[4];
        __syncthreads();
        if (tid < S) s_idx[tid] = base + tid < B ? (uint64_t)idx_in[base + tid] : 0ull;
        __syncthreads();
        // =========================== forward pass of the tile, keeping what the backward needs ===========================
        if (PHASE != 1) {
        if (!have_w) fetch_w_t(wv, P.in_proj_w[0], 0);
        for (int e = tid; e < 64 * 64; e += MD_THREADS) {
            const int k = e >> 6, r = e & 63;
            double v = 0.0;
            if (r < rows) {
                const int s = r / T, t = r - s * T;
                const int tok = t == 0 ? 2 : (int)((s_idx[s] >> (t - 1)) & 1ull);
                v = __ldg(P.tok_emb + tok * TB_D + k) + __ldg(P.pos_emb + t * TB_D + k);
            }
            X[k * MD_S + r] = v;
        }
        for (int l = 0; l < L; ++l) {
            __syncthreads();
            store_rows(X, ws.xin[l], row0, live);
            double *dst[3] = {Qb, Kb, Vb};
            for (int part = 0; part < 3; ++part) {
                __syncthreads();
                commit_w_t(wt, wv);
                if (part < 2) fetch_w_t(wv, P.in_proj_w[l], (part + 1) * TB_D); else fetch_w_t(wv, P.out_proj_w[l], 0);
                __syncthreads();
                double acc[4][4];
                gemm_tile(X, wt, TB_D, tx, ty, acc, dst[part]);
                store_acc(dst[part], acc, P.in_proj_b[l] ? P.in_proj_b[l] + part * TB_D : nullptr, tx, ty);
            }
            __syncthreads();
            store_rows(Qb, ws.q[l], row0, live);
            store_rows(Kb, ws.k[l], row0, live);
            store_rows(Vb, ws.v[l], row0, live);
            __syncthreads();
            for (int pair = tid; pair < rows * H; pair += MD_THREADS) {   // causal attention; the output overwrites the query slice
                const int h = pair / rows, r = pair - h * rows;            // lanes run over rows: K / V reads of a sample broadcast
                const int s = r / T, t = r - s * T, r0 = s * T;
                double q[HD], o[HD];
                _Pragma("unroll") for (int d = 0; d < hd; ++d) q[d] = Qb[(h * hd + d) * MD_S + r] * scale;
                double mx = -INFINITY;
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) sc += q[d] * Kb[(h * hd + d) * MD_S + r0 + tp];
                    mx = fmax(mx, sc);
                }
                double den = 0.0;
                _Pragma("unroll") for (int d = 0; d < hd; ++d) o[d] = 0.0;
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) sc += q[d] * Kb[(h * hd + d) * MD_S + r0 + tp];
                    const double p = exp(sc - mx);
                    den += p;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) o[d] += p * Vb[(h * hd + d) * MD_S + r0 + tp];
                }
                _Pragma("unroll") for (int d = 0; d < hd; ++d) Qb[(h * hd + d) * MD_S + r] = o[d] / den;
            }
            __syncthreads();
            store_rows(Qb, ws.a[l], row0, live);
            commit_w_t(wt, wv);
            fetch_w_t(wv, P.lin1_w[l], 0);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(Qb, wt, TB_D, tx, ty, acc, Kb);
                residual_layer_norm_save(X, Kb, acc, P.out_proj_b[l], P.ln1_w[l], P.ln1_b[l], P.ln_eps, tx, ty);
            }
            __syncthreads();
            store_rows(Kb, ws.y1[l], row0, live);
            store_rows(X, ws.x1[l], row0, live);
            commit_w_t(wt, wv);
            fetch_w_t(wv, P.lin2_w[l], 0);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(X, wt, TB_D, tx, ty, acc, Kb);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = tx + 16 * jj;
                    const double b = P.lin1_b[l] ? __ldg(P.lin1_b[l] + j) : 0.0;
                    double2 *o2 = reinterpret_cast<double2 *>(Kb + j * MD_S + ty * 4);
                    o2[0] = make_double2(fmax(acc[0][jj] + b, 0.0), fmax(acc[1][jj] + b, 0.0));
                    o2[1] = make_double2(fmax(acc[2][jj] + b, 0.0), fmax(acc[3][jj] + b, 0.0));
                }
            }
            __syncthreads();
            store_rows(Kb, ws.hf[l], row0, live);
            commit_w_t(wt, wv);
            have_w = false;
            if (l + 1 < L) {
                fetch_w_t(wv, P.in_proj_w[l + 1], 0);
            } else if (PHASE == 0 && tile + gridDim.x < ntiles) {
                fetch_w_t(wv, P.in_proj_w[0], 0);   // for this CTA's next tile
                have_w = true;
            }
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(Kb, wt, TB_D, tx, ty, acc, Qb);
                residual_layer_norm_save(X, Qb, acc, P.lin2_b[l], P.ln2_w[l], P.ln2_b[l], P.ln_eps, tx, ty);
            }
            __syncthreads();
            store_rows(Qb, ws.y2[l], row0, live);
        }
        store_rows(X, ws.xf, row0, live);
        {   // decoder: 4 numbers per token
            const int r = tid >> 2, c = tid & 3;
            double acc = __ldg(P.dec_b + c);
            for (int k = 0; k < TB_D; ++k) acc = fma(X[k * MD_S + r], __ldg(P.dec_w + c * TB_D + k), acc);
            s_dec[r * 4 + c] = acc;
        }
        __syncthreads();
        if (PHASE == 0) {   // log psi: one thread per position, then one per sample adds the positions up in order
            for (int e = tid; e < live * 4; e += MD_THREADS) ws.gdec[row0 * 4 + e] = s_dec[e];   // PHASE 1 starts from these
            if (tid < rows) {
                const int sm = tid / T, t = tid - sm * T;
                const uint64_t x = s_idx[sm];
                const double *o = s_dec + tid * 4;  // (re0, im0, re1, im1)
                const uint64_t prefix = t == 0 ? 0ull : (x & ((1ull << t) - 1ull));
                const long long mi = memo_index_of(P.sym_num, P.sym, prefix);
                const uint64_t mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)t * P.memo_size + mi) : 0ull;
                const bool a0 = mw & 1ull, a1 = (mw >> 1) & 1ull;
                const double z0 = a0 ? o[0] : -INFINITY, z1 = a1 ? o[2] : -INFINITY;
                const double mx = fmax(z0, z1);
                const double Ln = mx + 0.5 * log((a0 ? exp(2.0 * (z0 - mx)) : 0.0) + (a1 ? exp(2.0 * (z1 - mx)) : 0.0));
                const int bit = (int)((x >> t) & 1ull);
                const bool ok = bit ? a1 : a0;
                st_mx[tid] = ok ? (bit ? z1 : z0) - Ln : 0.0;
                st_den[tid] = ok ? (bit ? o[3] : o[1]) : 0.0;
                st_d[tid] = ok ? 0.0 : 1.0;
            }
            __syncthreads();
            if (tid < S && base + tid < B) {
                double re = 0.0, im = 0.0;
                bool dead = false;
                for (int t = 0; t < T; ++t) {
                    re += st_mx[tid * T + t];
                    im += st_den[tid * T + t];
                    dead |= st_d[tid * T + t] != 0.0;
                }
                log_psi[base + tid] = dead ? make_double2(-INFINITY, 0.0) : make_double2(re, im);
            }
        }
        }   // PHASE != 1
        if (PHASE == 0) continue;
        // =========================== gradient of log psi with respect to the decoder outputs ===============================
        // per position: re = z_bit - L, L = 0.5 logsumexp(2 z) over the allowed outcomes; im = the chosen outcome's phase
        if (PHASE == 1) {
            for (int e = tid; e < 64 * 4; e += MD_THREADS) s_dec[e] = e < live * 4 ? ws.gdec[row0 * 4 + e] : 0.0;
            __syncthreads();
        }
        if (tid < rows) {
            const int sm = tid / T, t = tid - sm * T;
            const uint64_t x = s_idx[sm];
            const double2 g = base + sm < B ? grad_out[base + sm] : make_double2(0.0, 0.0);
            double *o = s_dec + tid * 4;  // (re0, im0, re1, im1) -> their gradients
            const uint64_t prefix = t == 0 ? 0ull : (x & ((1ull << t) - 1ull));
            const long long mi = memo_index_of(P.sym_num, P.sym, prefix);
            const uint64_t mw = (mi >= 0 && mi < P.memo_size) ? __ldg(P.cont_mask + (size_t)t * P.memo_size + mi) : 0ull;
            const bool a0 = mw & 1ull, a1 = (mw >> 1) & 1ull;
            const double z0 = a0 ? o[0] : -INFINITY, z1 = a1 ? o[2] : -INFINITY;
            const double mx = fmax(z0, z1);
            const double e0 = a0 ? exp(2.0 * (z0 - mx)) : 0.0, e1 = a1 ? exp(2.0 * (z1 - mx)) : 0.0;
            const double p0 = (a0 || a1) ? e0 / (e0 + e1) : 0.0, p1 = (a0 || a1) ? e1 / (e0 + e1) : 0.0;
            const int bit = (int)((x >> t) & 1ull);
            const bool chosen_ok = bit ? a1 : a0;
            o[0] = a0 ? g.x * (((bit == 0 && chosen_ok) ? 1.0 : 0.0) - p0) : 0.0;
            o[2] = a1 ? g.x * (((bit == 1 && chosen_ok) ? 1.0 : 0.0) - p1) : 0.0;
            o[1] = bit == 0 ? g.y : 0.0;
            o[3] = bit == 1 ? g.y : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < live * 4; e += MD_THREADS) ws.gdec[row0 * 4 + e] = s_dec[e];
        // =========================== backward through the layers =============================================================
        double *G = buf[0], *B1 = buf[1], *B2 = buf[2], *B3 = buf[3], *W = buf[4];
        for (int e = tid; e < 64 * 64; e += MD_THREADS) {   // dX = dDec * W_dec
            const int k = e >> 6, r = e & 63;
            double v = 0.0;
            if (r < live)
                for (int c = 0; c < 4; ++c) v = fma(s_dec[r * 4 + c], __ldg(P.dec_w + c * TB_D + k), v);
            G[k * MD_S + r] = v;
        }
        fetch_w_n(wv, P.lin2_w[L - 1], 0);
        for (int l = L - 1; l >= 0; --l) {
            double *vl = vecs + (size_t)l * 4 * 64;
            __syncthreads();
            // ---- LayerNorm 2 ------------------------------------------------------------------------------------------------
            load_rows(B1, ws.y2[l], row0, live);
            __syncthreads();
            layer_norm_backward(G, B1, P.ln2_w[l], P.ln_eps, vl + 2 * 64, vl + 3 * 64, red, live, tx, ty);
            store_rows(G, ws.gy2[l], row0, live);                 // dy2: gradient of the second feed-forward linear's output
            // ---- feed-forward ---------------------------------------------------------------------------------------------
            commit_w_n(W, wv);
            fetch_w_n(wv, P.lin1_w[l], 0);
            load_rows(B2, ws.hf[l], row0, live);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(G, W, TB_D, tx, ty, acc, B1);           // dHf = dy2 * W2
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                {
                    const int a = (tx + 16 * jj) * MD_S + ty * 4;
                    const double2 h01 = *reinterpret_cast<const double2 *>(B2 + a), h23 = *reinterpret_cast<const double2 *>(B2 + a + 2);
                    *reinterpret_cast<double2 *>(B1 + a) = make_double2(h01.x > 0.0 ? acc[0][jj] : 0.0, h01.y > 0.0 ? acc[1][jj] : 0.0);   // ReLU
                    *reinterpret_cast<double2 *>(B1 + a + 2) = make_double2(h23.x > 0.0 ? acc[2][jj] : 0.0, h23.y > 0.0 ? acc[3][jj] : 0.0);
                }
            }
            __syncthreads();
            store_rows(B1, ws.ghp[l], row0, live);
            commit_w_n(W, wv);
            fetch_w_n(wv, P.out_proj_w[l], 0);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(B1, W, TB_D, tx, ty, acc, B1);          // dX1 = dy2 (residual) + dHpre * W1 (B1 is consumed)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    double2 *g2 = reinterpret_cast<double2 *>(G + (tx + 16 * jj) * MD_S + ty * 4);
                    const double2 u = g2[0], v = g2[1];
                    g2[0] = make_double2(u.x + acc[0][jj], u.y + acc[1][jj]);
                    g2[1] = make_double2(v.x + acc[2][jj], v.y + acc[3][jj]);
                }
            }
            __syncthreads();
            // ---- LayerNorm 1 ------------------------------------------------------------------------------------------------
            load_rows(B2, ws.y1[l], row0, live);
            __syncthreads();
            layer_norm_backward(G, B2, P.ln1_w[l], P.ln_eps, vl, vl + 64, red, live, tx, ty);
            store_rows(G, ws.gy1[l], row0, live);                 // dy1: gradient of the output projection's output (and of X_in)
            // ---- attention ------------------------------------------------------------------------------------------------
            commit_w_n(W, wv);
            fetch_w_n(wv, P.in_proj_w[l], 0);
            __syncthreads();
            {
                double acc[4][4];
                gemm_tile(G, W, TB_D, tx, ty, acc, B3);           // dA = dy1 * W_o
                store_acc(B3, acc, nullptr, tx, ty);
            }
            __syncthreads();
            // G's buffer is free from here (dy1 is in the workspace): Q -> B1, K -> B2, V -> G, dA in B3, dQ -> W
            load_rows(B1, ws.q[l], row0, live);
            load_rows(B2, ws.k[l], row0, live);
            load_rows(G, ws.v[l], row0, live);
            __syncthreads();
            for (int pair = tid; pair < rows * H; pair += MD_THREADS) {   // pass 1, per (head, query row): statistics and dQ
                const int h = pair / rows, r = pair - h * rows;
                const int s = r / T, t = r - s * T, r0 = s * T;
                double q[HD], da[HD], dq[HD];
                _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                    q[d] = B1[(h * hd + d) * MD_S + r] * scale;
                    da[d] = B3[(h * hd + d) * MD_S + r];
                    dq[d] = 0.0;
                }
                double mx = -INFINITY;
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) sc += q[d] * B2[(h * hd + d) * MD_S + r0 + tp];
                    mx = fmax(mx, sc);
                }
                double den = 0.0, dsum = 0.0;
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0, dp = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                        sc += q[d] * B2[(h * hd + d) * MD_S + r0 + tp];
                        dp += da[d] * G[(h * hd + d) * MD_S + r0 + tp];
                    }
                    const double p = exp(sc - mx);
                    den += p;
                    dsum += p * dp;
                }
                dsum /= den;   // sum_tp P dP
                for (int tp = 0; tp <= t; ++tp) {
                    double sc = 0.0, dp = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                        sc += q[d] * B2[(h * hd + d) * MD_S + r0 + tp];
                        dp += da[d] * G[(h * hd + d) * MD_S + r0 + tp];
                    }
                    const double ds = exp(sc - mx) / den * (dp - dsum);
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) dq[d] += ds * B2[(h * hd + d) * MD_S + r0 + tp];
                }
                _Pragma("unroll") for (int d = 0; d < hd; ++d) W[(h * hd + d) * MD_S + r] = dq[d] * scale;
                st_mx[pair] = mx;
                st_den[pair] = den;
                st_d[pair] = dsum;
            }
            // rows beyond the tile's samples: dQ = 0
            for (int e = tid; e < 64 * 64; e += MD_THREADS)
                if ((e & 63) >= rows) W[(e >> 6) * MD_S + (e & 63)] = 0.0;
            __syncthreads();
            for (int pair = tid; pair < rows * H; pair += MD_THREADS) {   // pass 2, per (head, key row): dK, dV in place of K, V
                const int h = pair / rows, rk = pair - h * rows;
                const int s = rk / T, tp = rk - s * T, r0 = s * T;
                double kk[HD], vv[HD], dk[HD], dv[HD];
                _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                    kk[d] = B2[(h * hd + d) * MD_S + rk];
                    vv[d] = G[(h * hd + d) * MD_S + rk];
                    dk[d] = dv[d] = 0.0;
                }
                for (int t = 0; t < T; ++t) {   // every lane walks the same query rows (broadcast reads); causal: t >= tp only
                    if (t < tp) continue;
                    const int r = r0 + t, pr = h * rows + r;
                    double sc = 0.0, dp = 0.0;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                        sc += B1[(h * hd + d) * MD_S + r] * kk[d];
                        dp += B3[(h * hd + d) * MD_S + r] * vv[d];
                    }
                    const double p = exp(sc * scale - st_mx[pr]) / st_den[pr];
                    const double ds = p * (dp - st_d[pr]) * scale;
                    _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                        dk[d] += ds * B1[(h * hd + d) * MD_S + r];
                        dv[d] += p * B3[(h * hd + d) * MD_S + r];
                    }
                }
                _Pragma("unroll") for (int d = 0; d < hd; ++d) {
                    B2[(h * hd + d) * MD_S + rk] = dk[d];   // only this thread reads these entries of K and V in this pass
                    G[(h * hd + d) * MD_S + rk] = dv[d];
                }
            }
            __syncthreads();
            store_rows(W, ws.gqkv[l], row0, live, 192, 0);
            store_rows(B2, ws.gqkv[l], row0, live, 192, 64);
            store_rows(G, ws.gqkv[l], row0, live, 192, 128);
            // ---- dX_in = dy1 (residual) + dQ W_q + dK W_k + dV W_v ---------------------------------------------------------
            load_rows(B1, ws.gy1[l], row0, live);
            double tot[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) tot[a][b] = 0.0;
            double *src[3] = {W, B2, G};
            for (int part = 0; part < 3; ++part) {
                __syncthreads();
                commit_w_n(B3, wv);
                if (part < 2) fetch_w_n(wv, P.in_proj_w[l], (part + 1) * TB_D); else if (l > 0) fetch_w_n(wv, P.lin2_w[l - 1], 0);
                __syncthreads();
                double acc[4][4];
                gemm_tile(src[part], B3, TB_D, tx, ty, acc, src[part]);   // dQ / dK / dV are consumed
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) tot[a][b] += acc[a][b];
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                double2 *g2 = reinterpret_cast<double2 *>(B1 + (tx + 16 * jj) * MD_S + ty * 4);
                const double2 u = g2[0], v = g2[1];
                g2[0] = make_double2(u.x + tot[0][jj], u.y + tot[1][jj]);
                g2[1] = make_double2(v.x + tot[2][jj], v.y + tot[3][jj]);
            }
            // the running gradient now lives in B1
            double *tmp = G;
            G = B1;
            B1 = tmp;
        }
        __syncthreads();
        // ---- embeddings: dX0 per row and the one-hot (token | position) row that batch_reduce multiplies it with -----------------
        store_rows(G, ws.gx0, row0, live);
        for (int e = tid; e < live * ws.P; e += MD_THREADS) {
            const int r = e / ws.P, c = e - r * ws.P;
            const int s = r / T, t = r - s * T;
            const int tok = t == 0 ? 2 : (int)((s_idx[s] >> (t - 1)) & 1ull);
            ws.emb[row0 * ws.P + e] = (c == tok || c == 3 + t) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    if (PHASE != 0)
        for (int e = tid; e < TB_NVEC * 64; e += MD_THREADS) ws.vec[(size_t)blockIdx.x * TB_NVEC * 64 + e] = vecs[e];
}

// LayerNorm weight / bias gradients: sum of the per-CTA partial sums, in CTA order
__global__ void transformer_vec_finish_kernel(const double *__restrict__ vec, int n_cta, int accumulate, int depth, anqs_transformer_grads_t g) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= depth * 4 * 64) return;
    const int l = e >> 8, which = (e >> 6) & 3, j = e & 63;
    double *dst = which == 0 ? g.ln1_w[l] : which == 1 ? g.ln1_b[l] : which == 2 ? g.ln2_w[l] : g.ln2_b[l];
    if (!dst) return;
    double s = accumulate ? dst[j] : 0.0;
    for (int c = 0; c < n_cta; ++c) s += vec[(size_t)c * TB_NVEC * 64 + e];
    dst[j] = s;
}

static int64_t tb_row_doubles(const anqs_transformer_desc_t *desc) {
    return (int64_t)desc->depth * (9 * 64 + 192 + 3 * 64) + 64 + 4 + 64 + 3 + desc->qubit_num;
}

// carves the per-row workspace and lists the batch reductions (grad W = dY^T X, grad b = column sums of dY)
static int tb_layout(const anqs_transformer_desc_t *desc, const anqs_transformer_grads_t *g, double *base, int64_t n, TfBwdPtrs &ws,
                     anqs_brg_problem_t *problems) {
    const int T = desc->qubit_num, L = desc->depth, P = 3 + T;
    const int64_t R = n * T;
    double *p = base;
    auto take = [&](int64_t per_row) { double *q = p; p += R * per_row; return q; };
    for (int l = 0; l < 4; ++l) {
        const bool on = l < L;
        ws.xin[l] = on ? take(64) : nullptr; ws.q[l] = on ? take(64) : nullptr; ws.k[l] = on ? take(64) : nullptr;
        ws.v[l] = on ? take(64) : nullptr; ws.a[l] = on ? take(64) : nullptr; ws.y1[l] = on ? take(64) : nullptr;
        ws.x1[l] = on ? take(64) : nullptr; ws.hf[l] = on ? take(64) : nullptr; ws.y2[l] = on ? take(64) : nullptr;
        ws.gqkv[l] = on ? take(192) : nullptr; ws.gy1[l] = on ? take(64) : nullptr; ws.ghp[l] = on ? take(64) : nullptr;
        ws.gy2[l] = on ? take(64) : nullptr;
    }
    ws.xf = take(64);
    ws.gdec = take(4);
    ws.gx0 = take(64);
    ws.emb = take(P);
    ws.vec = p;
    ws.P = P;
    if (!g || !problems) return 0;
    int np = 0;
    auto add = [&](const double *A, int lda, int M, const double *B, double *C, double *colsum) {
        anqs_brg_problem_t q;
        q.A = A, q.B = B, q.C = C, q.colsum = colsum;
        q.lda = lda, q.ldb = 64, q.ldc = 64, q.M = M, q.N = 64, q.reserved = 0;
        problems[np++] = q;
    };
    for (int l = 0; l < L; ++l) {
        add(ws.gqkv[l], 192, 192, ws.xin[l], g->in_proj_w[l], g->in_proj_b[l]);
        add(ws.gy1[l], 64, 64, ws.a[l], g->out_proj_w[l], g->out_proj_b[l]);
        add(ws.ghp[l], 64, 64, ws.x1[l], g->lin1_w[l], g->lin1_b[l]);
        add(ws.gy2[l], 64, 64, ws.hf[l], g->lin2_w[l], g->lin2_b[l]);
    }
    add(ws.gdec, 4, 4, ws.xf, g->dec_w, g->dec_b);
    add(ws.emb, P, 3, ws.gx0, g->tok_emb, nullptr);
    add(ws.emb + 3, P, T, ws.gx0, g->pos_emb, nullptr);
    return np;
}

}  // namespace anqs

using namespace anqs;

static int tb_check_desc(const anqs_transformer_desc_t *desc) {
    ANQS_REQUIRE(desc, "null descriptor");
    ANQS_REQUIRE(desc->dim == TB_D && desc->depth >= 1 && desc->depth <= 4, "model dimension must be 64, depth 1..4");
    ANQS_REQUIRE(desc->head_num == 1 || desc->head_num == 2 || desc->head_num == 4 || desc->head_num == 8 || desc->head_num == 16,
                 "head_num must be 1, 2, 4, 8 or 16");
    ANQS_REQUIRE(desc->qubit_num >= 1 && desc->qubit_num <= 64, "qubit_num must be in [1, 64]");
    return 0;
}
static int tb_check(const anqs_transformer_desc_t *desc, const anqs_transformer_grads_t *g) {
    if (int rc = tb_check_desc(desc)) return rc;
    ANQS_REQUIRE(g, "null gradient descriptor");
    ANQS_REQUIRE(g->tok_emb && g->pos_emb && g->dec_w && g->dec_b, "null embedding / decoder gradient pointer");
    for (int l = 0; l < desc->depth; ++l)
        ANQS_REQUIRE(g->in_proj_w[l] && g->out_proj_w[l] && g->lin1_w[l] && g->lin2_w[l] && g->ln1_w[l] && g->ln1_b[l] && g->ln2_w[l] && g->ln2_b[l],
                     "null weight gradient pointer");
    return 0;
}
static int64_t tb_own_bytes(const anqs_transformer_desc_t *desc, int64_t n) {
    return (n * desc->qubit_num * tb_row_doubles(desc) + (int64_t)sm_count_of_current_device() * TB_NVEC * 64) * 8;
}

template <int PHASE>
static int tb_launch(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out, double *d_log_psi,
                     const TfBwdPtrs &ws, cudaStream_t s) {
    const int T = desc->qubit_num;
    const int grid = (int)std::min<int64_t>((n + (64 / T) - 1) / (64 / T), (int64_t)sm_count_of_current_device());
    auto launch = [&](auto kern) -> int {
        ANQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM));
        kern<<<grid, MD_THREADS, TB_SMEM, s>>>(*desc, d_idx, n, (const double2 *)d_grad_out, (double2 *)d_log_psi, ws);
        ANQS_LAUNCH_CHECK();
        return 0;
    };
    switch (desc->head_num) {
        case 1: return launch(transformer_backward_kernel<64, PHASE>);
        case 2: return launch(transformer_backward_kernel<32, PHASE>);
        case 4: return launch(transformer_backward_kernel<16, PHASE>);
        case 8: return launch(transformer_backward_kernel<8, PHASE>);
        default: return launch(transformer_backward_kernel<4, PHASE>);
    }
}

extern "C" {

int64_t anqs_transformer_backward_workspace(const anqs_transformer_desc_t *desc, int64_t n) {
    if (!desc || n <= 0 || tb_check_desc(desc)) return -1;
    static double dummy[1];
    anqs_transformer_grads_t g;
    g.tok_emb = g.pos_emb = g.dec_w = g.dec_b = dummy;
    for (int l = 0; l < 4; ++l)
        g.in_proj_w[l] = g.in_proj_b[l] = g.out_proj_w[l] = g.out_proj_b[l] = g.lin1_w[l] = g.lin1_b[l] = g.lin2_w[l] = g.lin2_b[l] = g.ln1_w[l] =
            g.ln1_b[l] = g.ln2_w[l] = g.ln2_b[l] = dummy;
    TfBwdPtrs ws;
    anqs_brg_problem_t problems[4 * 4 + 3];
    const int np = tb_layout(desc, &g, dummy, n, ws, problems);   // addresses are not dereferenced here
    const int64_t brg = anqs_batch_reduce_workspace(problems, np, n * desc->qubit_num);
    if (brg < 0) return -1;
    return tb_own_bytes(desc, n) + 256 + brg;
}

int anqs_transformer_log_psi_saving(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, void *d_work,
                                    int64_t work_bytes, void *stream) {
    if (int rc = tb_check_desc(desc)) return rc;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_log_psi && d_work, "null pointer");
    ANQS_REQUIRE(work_bytes >= anqs_transformer_backward_workspace(desc, n), "workspace too small");
    TfBwdPtrs ws;
    tb_layout(desc, nullptr, (double *)d_work, n, ws, nullptr);
    return tb_launch<0>(desc, d_idx, n, nullptr, d_log_psi, ws, (cudaStream_t)stream);
}

int anqs_transformer_backward(const anqs_transformer_desc_t *desc, const anqs_transformer_grads_t *grads, const int64_t *d_idx, int64_t n,
                              const double *d_grad_out, void *d_work, int64_t work_bytes, int saved, int accumulate, void *stream) {
    if (int rc = tb_check(desc, grads)) return rc;
    ANQS_REQUIRE(n >= 0, "negative sample count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_idx && d_grad_out && d_work, "null pointer");
    ANQS_REQUIRE(work_bytes >= anqs_transformer_backward_workspace(desc, n), "workspace too small");
    const int T = desc->qubit_num;
    const int grid = (int)std::min<int64_t>((n + (64 / T) - 1) / (64 / T), (int64_t)sm_count_of_current_device());
    TfBwdPtrs ws;
    anqs_brg_problem_t problems[4 * 4 + 3];
    const int np = tb_layout(desc, grads, (double *)d_work, n, ws, problems);
    const int64_t own = tb_own_bytes(desc, n);
    char *brg_ws = (char *)d_work + ((own + 255) / 256) * 256;
    const int64_t brg_bytes = work_bytes - (brg_ws - (char *)d_work);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = saved ? tb_launch<1>(desc, d_idx, n, d_grad_out, nullptr, ws, s) : tb_launch<2>(desc, d_idx, n, d_grad_out, nullptr, ws, s)) return rc;
    transformer_vec_finish_kernel<<<(desc->depth * 4 * 64 + 255) / 256, 256, 0, s>>>(ws.vec, grid, accumulate, desc->depth, *grads);
    ANQS_LAUNCH_CHECK();
    if (!accumulate)   // the last positional row (position qubit_num) never enters the network
        ANQS_CUDA(cudaMemsetAsync(grads->pos_emb + (size_t)T * TB_D, 0, TB_D * sizeof(double), s));
    return anqs_batch_reduce_gemm(problems, np, n * T, accumulate, brg_ws, brg_bytes, stream);
}

}  // extern "C"
