// Matrix elements H_{x,x'} = sum over the YZ group of weight * (-1)^popcount(x' & yz)  (reference PO:256-324),
// shared by the emit, matrix-element and fused local-energy kernels.
#pragma once
#include "common.cuh"

namespace anqs {

constexpr int BIG_GROUP = 48;  // YZ groups longer than this are summed by the whole warp

// ---- H_{x,x'} = sum_t w_t (-1)^{popcount(x' & yz_t)} over one YZ group, per lane ------------------
// xp and the YZ masks are both DE-INTERLEAVED (a bit permutation, so the parity is unchanged).
template <bool REAL>
__device__ __forceinline__ void group_sum_lane(const Tables &t, int start, int num, uint64_t xp, double &hr,
                                               double &hi) {
    hr = 0.0;
    hi = 0.0;
    if (REAL) {
        const ulonglong2 *rec = t.term_real + start;
#pragma unroll 4
        for (int k = 0; k < num; ++k) {
            ulonglong2 r = __ldg(rec + k);
            hr += flip_sign(__longlong_as_double((long long)r.y), parity64(xp & r.x));
        }
    } else {
        for (int k = start; k < start + num; ++k) {
            uint32_t par = parity64(xp & __ldg(t.yz_d + k));
            hr += flip_sign(__ldg(t.w_re + k), par);
            hi += flip_sign(__ldg(t.w_im + k), par);
        }
    }
}

// whole warp sums one group; every lane returns the total
template <bool REAL>
__device__ __forceinline__ void group_sum_warp(const Tables &t, int start, int num, uint64_t xp, double &hr,
                                               double &hi) {
    double sr = 0.0, si = 0.0;
    for (int k = start + lane_id(); k < start + num; k += 32) {
        if (REAL) {
            ulonglong2 r = __ldg(t.term_real + k);
            sr += flip_sign(__longlong_as_double((long long)r.y), parity64(xp & r.x));
        } else {
            uint32_t par = parity64(xp & __ldg(t.yz_d + k));
            sr += flip_sign(__ldg(t.w_re + k), par);
            si += flip_sign(__ldg(t.w_im + k), par);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, d);
        if (!REAL) si += __shfl_xor_sync(0xffffffffu, si, d);
    }
    hr = sr;
    hi = si;
}

// Matrix elements for up to 32 (group, x') pairs held one per lane.  Must be called by the whole warp.
template <bool REAL>
__device__ __forceinline__ void warp_matrix_elements(const Tables &t, bool active, int2 g, uint64_t xp, double &hr,
                                                     double &hi) {
    hr = 0.0;
    hi = 0.0;
    bool big = active && g.y > BIG_GROUP;
    if (active && !big) group_sum_lane<REAL>(t, g.x, g.y, xp, hr, hi);
    unsigned bigmask = __ballot_sync(0xffffffffu, big);
    while (bigmask) {
        int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        int gx = __shfl_sync(0xffffffffu, g.x, src);
        int gy = __shfl_sync(0xffffffffu, g.y, src);
        uint64_t xs = __shfl_sync(0xffffffffu, xp, src);
        double sr, si;
        group_sum_warp<REAL>(t, gx, gy, xs, sr, si);
        if (lane_id() == src) {
            hr = sr;
            hi = si;
        }
    }
}

}  // namespace anqs
