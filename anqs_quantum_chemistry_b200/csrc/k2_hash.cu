// Kernel family 2: membership join of connected configurations against the sampled set
// (reference HS:263-284 find_a_in_b: cat -> unique(return_inverse) -> scatter_ -> gather).
//
// Open-addressing table with linear probing.  One slot is one 32-byte sector
// {key u64 (de-interleaved), index i64, amp.re f64, amp.im f64}, so a probe that hits returns the amplitude
// psi(x') from the same sector.  A blocked Bloom filter (capacity/4 words of 32 bits behind the slots, 3
// bits of one word per key, >= 16 bits per key) decides ~99 % of the misses with one 4-byte load of an
// array that stays L2-resident even when the slot array does not.  The all-ones key is the EMPTY sentinel; a real all-ones key (only possible at
// qubit_num == 64, or for generic int64 inputs such as -1) lives in a dedicated slot at index `capacity`.
#include <algorithm>

#include "common.cuh"

namespace anqs {

__global__ void __launch_bounds__(256)
hash_build_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                  uint32_t *bloom, uint32_t capmask, uint32_t wordmask) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        uint64_t key = deinterleave((uint64_t)keys[j]);
        HashSlot *sl;
        if (key == EMPTY_KEY) {
            sl = slots + (size_t)capmask + 1;
        } else {
            uint32_t hh = hash_key((uint32_t)key, (uint32_t)(key >> 32));
            atomicOr(bloom + (bloom_word(hh) & wordmask), bloom_pattern(hh));
            uint32_t h = hh & capmask;
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
hash_probe_kernel(HashView hv, const int64_t *__restrict__ queries, int64_t m, int64_t *__restrict__ ptr,
                  uint8_t *__restrict__ mask) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        uint64_t q = deinterleave((uint64_t)queries[i]);
        long long j = -1;
        bool maybe = true;
        if (q != EMPTY_KEY) {
            uint32_t hh = hash_key((uint32_t)q, (uint32_t)(q >> 32)), pat = bloom_pattern(hh);
            maybe = (__ldg(hv.bloom + (bloom_word(hh) & hv.wordmask)) & pat) == pat;
        }
        if (maybe) {
            double re, im;
            j = hash_lookup(hv, q, re, im);
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

size_t anqs_hash_bytes(int64_t capacity) { return (size_t)(capacity + 1) * sizeof(HashSlot) + (size_t)capacity; }

int anqs_hash_build(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                    void *stream) {
    ANQS_REQUIRE(n >= 0, "negative key count");
    ANQS_REQUIRE(d_table, "null table buffer");
    ANQS_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 1024");
    ANQS_REQUIRE(capacity >= 2 * n, "capacity must be at least 2n (use anqs_hash_capacity)");
    ANQS_REQUIRE(capacity <= ((int64_t)1 << 29), "capacity above 2^29 slots is not supported");
    ANQS_REQUIRE(((uintptr_t)d_table & 31) == 0, "table buffer must be 32-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    HashSlot *slots = (HashSlot *)d_table;
    uint32_t *bits = (uint32_t *)(slots + capacity + 1);
    // 0xFF over the slots: key = EMPTY, idx = -1; zero Bloom words
    ANQS_CUDA(cudaMemsetAsync(slots, 0xFF, (size_t)(capacity + 1) * sizeof(HashSlot), s));
    ANQS_CUDA(cudaMemsetAsync(bits, 0, (size_t)capacity, s));
    if (n == 0) return 0;
    ANQS_REQUIRE(d_keys, "null key array");
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    hash_build_kernel<<<grid, 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, slots, bits, (uint32_t)(capacity - 1),
                                           (uint32_t)(capacity / 4 - 1));
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_hash_probe(const void *d_table, int64_t capacity, const int64_t *d_queries, int64_t m, int64_t *d_ptr,
                    uint8_t *d_mask, void *stream) {
    ANQS_REQUIRE(m >= 0, "negative query count");
    if (m == 0) return 0;
    ANQS_REQUIRE(d_table && d_queries, "null pointer");
    ANQS_REQUIRE(d_ptr || d_mask, "nothing to compute: both outputs are NULL");
    ANQS_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 1024");
    HashView hv = make_hash_view(d_table, capacity);
    int grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    hash_probe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hv, d_queries, m, d_ptr, d_mask);
    ANQS_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
