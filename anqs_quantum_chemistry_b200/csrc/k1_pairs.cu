// Kernel family 1, pair-join variant of the sample-aware local energy (reference: the 'trie' and
// 'all_to_all' coupling methods, PO:602-696 with utils/trie.py:8-125, which find the coupled pairs among the SAMPLED
// configurations themselves instead of enumerating all U connected configurations of every sample):
//
//   E_loc[i] = ( sum_{j in the sampled set, x_i ^ x_j an XY mask of H} H_{x_i,x_j} psi_j ) / psi_i
//
// Cost N^2 pair tests (one XOR + one POPC each) instead of N x U filter tests: the better algorithm whenever the sampled set is
// smaller than the Hamiltonian's mask list, the regime of the reference's notebook (N_unq = 1e4, NB:248-250), and independent
// of U.  The reference prunes the pairs with a trie over the bit strings on the CPU; on the GPU the pruning IS the popcount
// test - a pair survives when popcount(x_i ^ x_j) <= the largest mask weight - and the survivors (a handful per row) look their
// mask up in a hash table of the unique XY masks (the same open-addressing table the sampled set uses) and sum the group's
// terms from the global arrays.
//
// Layout: a CTA owns 128 rows (one per thread, x_i in registers) and a slice of the columns, staged 512 at a time in shared
// memory as (x_j, psi_j, in-sector flag): every thread reads the same column at the same time (broadcast, conflict-free).
// Column slices go to gridDim.y so that small batches still fill the machine; partial sums land in a workspace and are added
// in a fixed order (deterministic).
#include <algorithm>

#include "common.cuh"
#include "matrix_elements.cuh"

namespace anqs {

constexpr int PJ_ROWS = 128, PJ_COLS = 512, PJ_WARPS = PJ_ROWS / 32;
constexpr int PJ_LIST = 64;   // pending coupled pairs per warp (flushed at 32)

struct PjHit {   // a coupled pair waiting for its matrix element
    uint64_t xj;
    int32_t u;       // index of the XY mask x_i ^ x_j
    int16_t row;     // lane that owns row i
    int16_t col;     // column slot in the staged chunk (its amplitude)
};

// Matrix elements of up to 32 pending pairs of one warp, one per lane, and their accumulation into the owning lanes (in list
// order, which is column order for every row: deterministic).
template <bool REAL>
__device__ __forceinline__ void pj_flush(const Tables &t, PjHit *list, double2 *contrib, int cnt, const double2 *s_a, double &er, double &ei) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < cnt;
    PjHit h = list[lane];
    double hr = 0.0, hi = 0.0;
    int2 g = make_int2(0, 0);
    if (active) g = __ldg(t.grp + h.u);
    warp_matrix_elements<REAL>(t, active, g, deinterleave(h.xj), hr, hi);
    if (active) {
        const double2 a = s_a[h.col];
        contrib[lane] = make_double2(hr * a.x - hi * a.y, hr * a.y + hi * a.x);
    }
    __syncwarp();
    for (int e = 0; e < cnt; ++e) {
        if (list[e].row == lane) {
            er += contrib[e].x;
            ei += contrib[e].y;
        }
    }
    __syncwarp();
}

template <bool REAL>
__global__ void __launch_bounds__(PJ_ROWS) pair_join_kernel(Tables t, HashView masks, const int64_t *__restrict__ samples,
                                                           const double2 *__restrict__ amps, int64_t n_total, int64_t row_start, int64_t row_len,
                                                           int alpha, int beta, int max_weight, int64_t cols_per_slice,
                                                           double2 *__restrict__ partial) {
    __shared__ uint64_t s_x[PJ_COLS];
    __shared__ double2 s_a[PJ_COLS];
    __shared__ uint8_t s_ok[PJ_COLS];
    __shared__ PjHit s_list[PJ_WARPS][PJ_LIST];
    __shared__ double2 s_contrib[PJ_WARPS][32];
    __shared__ int s_cnt[PJ_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * PJ_ROWS + threadIdx.x;
    const bool have = r < row_len;
    const uint64_t xi = have ? (uint64_t)samples[row_start + r] : 0ull;
    const int64_t c_lo = (int64_t)blockIdx.y * cols_per_slice, c_hi = min(n_total, c_lo + cols_per_slice);
    double er = 0.0, ei = 0.0;
    if (lane == 0) s_cnt[warp] = 0;
    for (int64_t c0 = c_lo; c0 < c_hi; c0 += PJ_COLS) {
        const int cnt = (int)min((int64_t)PJ_COLS, c_hi - c0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += PJ_ROWS) {
            const uint64_t xj = (uint64_t)samples[c0 + k];
            // columns outside the (N_alpha, N_beta) sector never couple (the 'ham' path drops them with its electron-count filter)
            s_x[k] = xj;
            s_ok[k] = (__popcll(xj & 0x5555555555555555ULL) == alpha && __popcll(xj & 0xAAAAAAAAAAAAAAAAULL) == beta) ? 1 : 0;
            s_a[k] = amps[c0 + k];
        }
        __syncthreads();
        // every lane tests its row against the same column (broadcast reads); the rare survivors look their mask up and go onto
        // the warp's pending list, which is evaluated 32 pairs at a time, one per lane - so a hit costs the warp one lane's
        // work, not a divergent excursion of the whole warp
        for (int k = 0; k < cnt; ++k) {
            const uint64_t xj = s_x[k];
            const uint64_t m = xi ^ xj;
            if (have && __popcll(m) <= max_weight && s_ok[k]) {
                double dr, di;
                const long long u = hash_lookup(masks, deinterleave(m), dr, di);
                if (u >= 0) {
                    const int pos = atomicAdd(&s_cnt[warp], 1);
                    PjHit h;
                    h.xj = xj;
                    h.u = (int32_t)u;
                    h.row = (int16_t)lane;
                    h.col = (int16_t)k;
                    s_list[warp][pos] = h;
                }
            }
            __syncwarp();
            const int pending = *reinterpret_cast<volatile int *>(&s_cnt[warp]);
            if (pending >= 32) {
                pj_flush<REAL>(t, s_list[warp], s_contrib[warp], 32, s_a, er, ei);
                if (lane < pending - 32) s_list[warp][lane] = s_list[warp][32 + lane];
                __syncwarp();
                if (lane == 0) s_cnt[warp] = pending - 32;
                __syncwarp();
            }
        }
        // the amplitudes of this chunk go away with it: evaluate what is pending
        const int pending = *reinterpret_cast<volatile int *>(&s_cnt[warp]);
        if (pending > 0) {
            pj_flush<REAL>(t, s_list[warp], s_contrib[warp], pending, s_a, er, ei);
            if (lane == 0) s_cnt[warp] = 0;
            __syncwarp();
        }
    }
    if (have) partial[(size_t)blockIdx.y * row_len + r] = make_double2(er, ei);
}

__global__ void pair_join_finish_kernel(const double2 *__restrict__ partial, int slices, const double2 *__restrict__ amps, int64_t row_start,
                                        int64_t row_len, double2 *__restrict__ eloc) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < row_len; r += (int64_t)gridDim.x * blockDim.x) {
        double er = 0.0, ei = 0.0;
        for (int s = 0; s < slices; ++s) {
            const double2 p = partial[(size_t)s * row_len + r];
            er += p.x;
            ei += p.y;
        }
        const double2 a = amps[row_start + r];
        const double den = a.x * a.x + a.y * a.y;
        eloc[r] = make_double2((er * a.x + ei * a.y) / den, (ei * a.x - er * a.y) / den);
    }
}

// Column slices: a function of the size of the sampled set ONLY, so that a window of rows (the multi-GPU shard path) adds the
// same partial sums in the same order as the full evaluation and comes out bit-identical.
static int pj_slices(int64_t row_len, int64_t n_total) {
    (void)row_len;
    return (int)std::max<int64_t>(1, std::min<int64_t>(64, (n_total + 63) / 64));
}

}  // namespace anqs

using namespace anqs;

extern "C" {

size_t anqs_pair_join_workspace(int64_t row_len, int64_t n_total) {
    if (row_len <= 0 || n_total <= 0) return 0;
    return (size_t)pj_slices(row_len, n_total) * (size_t)row_len * sizeof(double2);
}

int anqs_local_energy_pair_join(const anqs_tables_t *h, const int64_t *d_samples, const double *d_amps, int64_t n_total,
                                int64_t row_start, int64_t row_len, const void *d_mask_table, int64_t mask_capacity, int alpha_num,
                                int beta_num, double *d_eloc, void *d_work, void *stream) {
    ANQS_REQUIRE(h, "null tables handle");
    ANQS_REQUIRE(row_start >= 0 && row_len >= 0 && row_start + row_len <= n_total, "row window out of range");
    if (row_len == 0) return 0;
    ANQS_REQUIRE(d_samples && d_amps && d_mask_table && d_eloc && d_work, "null pointer");
    ANQS_REQUIRE(mask_capacity >= 1024 && (mask_capacity & (mask_capacity - 1)) == 0, "mask_capacity must be a power of two >= 1024");
    const Tables *t = (const Tables *)h;
    ANQS_REQUIRE(mask_capacity >= 2 * t->U, "the mask table must be built over the U unique XY masks (anqs_hash_build)");
    cudaStream_t s = (cudaStream_t)stream;
    HashView masks = make_hash_view(d_mask_table, mask_capacity);
    const int slices = pj_slices(row_len, n_total);
    const int64_t cols = ((n_total + slices - 1) / slices + 31) / 32 * 32;
    dim3 grid((unsigned)((row_len + PJ_ROWS - 1) / PJ_ROWS), (unsigned)slices);
    if (t->weights_real)
        pair_join_kernel<true><<<grid, PJ_ROWS, 0, s>>>(*t, masks, d_samples, (const double2 *)d_amps, n_total, row_start, row_len, alpha_num,
                                                        beta_num, t->max_xy_weight, cols, (double2 *)d_work);
    else
        pair_join_kernel<false><<<grid, PJ_ROWS, 0, s>>>(*t, masks, d_samples, (const double2 *)d_amps, n_total, row_start, row_len, alpha_num,
                                                         beta_num, t->max_xy_weight, cols, (double2 *)d_work);
    ANQS_LAUNCH_CHECK();
    pair_join_finish_kernel<<<(unsigned)std::min<int64_t>((row_len + 255) / 256, 1024), 256, 0, s>>>((const double2 *)d_work, slices,
                                                                                                    (const double2 *)d_amps, row_start, row_len,
                                                                                                    (double2 *)d_eloc);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
