// Kernel family 2: membership join of connected configurations against the sampled set
// (reference HS:263-284 find_a_in_b: cat -> unique(return_inverse) -> scatter_ -> gather).
//
// Open-addressing table with linear probing.  One slot is one 32-byte sector
// {key u64, index i64, amp.re f64, amp.im f64}, so a probe that hits returns the amplitude psi(x')
// from the same sector.  The all-ones key is the EMPTY sentinel; a real all-ones key (only possible at
// qubit_num == 64, or for generic int64 inputs such as -1) lives in a dedicated slot at index `capacity`.
#include <algorithm>

#include "common.cuh"

namespace anqs {

constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;
struct __align__(32) HashSlot {
    uint64_t key;
    long long idx;
    double re, im;
};

__global__ void __launch_bounds__(256)
hash_build_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                  uint64_t capmask) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        uint64_t key = (uint64_t)keys[j];
        HashSlot *sl;
        if (key == EMPTY_KEY) {
            sl = slots + capmask + 1;
        } else {
            uint64_t h = mix64(key) & capmask;
            for (;;) {
                unsigned long long prev = atomicCAS((unsigned long long *)&slots[h].key, (unsigned long long)EMPTY_KEY,
                                                    (unsigned long long)key);
                if (prev == EMPTY_KEY || prev == key) break;
                h = (h + 1) & capmask;
            }
            sl = slots + h;
        }
        // duplicates: the largest position wins, which is what a sequential scatter_ leaves behind
        long long old = atomicMax(&sl->idx, (long long)j);
        if (old < j && amps) {
            double2 a = amps[j];
            sl->re = a.x;
            sl->im = a.y;
        }
    }
}

__global__ void __launch_bounds__(256)
hash_probe_kernel(const HashSlot *__restrict__ slots, uint64_t capmask, const int64_t *__restrict__ queries, int64_t m,
                  int64_t *__restrict__ ptr, uint8_t *__restrict__ mask) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        uint64_t q = (uint64_t)queries[i];
        long long j = -1;
        if (q == EMPTY_KEY) {
            j = slots[capmask + 1].idx;
        } else {
            uint64_t h = mix64(q) & capmask;
            for (;;) {
                ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2 *>(slots + h));
                if (kv.x == q) {
                    j = (long long)kv.y;
                    break;
                }
                if (kv.x == EMPTY_KEY) break;
                h = (h + 1) & capmask;
            }
        }
        if (ptr) ptr[i] = j;
        if (mask) mask[i] = j >= 0 ? 1 : 0;
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

int64_t anqs_hash_capacity(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

int anqs_hash_build(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_slots, int64_t capacity,
                    void *stream) {
    ANQS_REQUIRE(n >= 0, "negative key count");
    ANQS_REQUIRE(d_slots, "null slot array");
    ANQS_REQUIRE(capacity >= 2 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
    ANQS_REQUIRE(capacity >= 2 * n, "capacity must be at least 2n (use anqs_hash_capacity)");
    cudaStream_t s = (cudaStream_t)stream;
    // 0xFF everywhere: key = EMPTY, idx = -1 (capacity + 1 slots: the last one is the all-ones slot)
    ANQS_CUDA(cudaMemsetAsync(d_slots, 0xFF, (size_t)(capacity + 1) * sizeof(HashSlot), s));
    if (n == 0) return 0;
    ANQS_REQUIRE(d_keys, "null key array");
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    hash_build_kernel<<<grid, 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, (HashSlot *)d_slots, (uint64_t)capacity - 1);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_hash_probe(const void *d_slots, int64_t capacity, const int64_t *d_queries, int64_t m, int64_t *d_ptr,
                    uint8_t *d_mask, void *stream) {
    ANQS_REQUIRE(m >= 0, "negative query count");
    if (m == 0) return 0;
    ANQS_REQUIRE(d_slots && d_queries, "null pointer");
    ANQS_REQUIRE(d_ptr || d_mask, "nothing to compute: both outputs are NULL");
    ANQS_REQUIRE(capacity >= 2 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
    int grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    hash_probe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const HashSlot *)d_slots, (uint64_t)capacity - 1, d_queries,
                                                              m, d_ptr, d_mask);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
