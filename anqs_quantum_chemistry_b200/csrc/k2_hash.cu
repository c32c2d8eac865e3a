// Kernel family 2: membership join of connected configurations against the sampled set
// (reference HS:263-284 find_a_in_b: cat -> unique(return_inverse) -> scatter_ -> gather).
//
// Open-addressing table with linear probing.  One slot is one 32-byte sector
// {key u64 (de-interleaved), index i64, amp.re f64, amp.im f64}, so a probe that hits returns the amplitude
// psi(x') from the same sector.  In front of it sits a line-blocked presence filter (layout and rationale in
// common.cuh): one bit per key, 64..128 bits per key, the 128-byte line chosen by a GF(2)-linear hash of the alpha
// half of the key.  The fused local-energy kernel (k1_fused.cu) consults only the filter for ~97 % of its
// candidates.  The all-ones key is the EMPTY sentinel; a real all-ones key (only possible at qubit_num == 64, or for
// generic int64 inputs such as -1) lives in a dedicated slot at index `capacity`.
//
// Build = stream-ordered launches, no host synchronisation:
//   memset -> count keys per line (in the filter region itself) -> pick the spread G -> memset filter -> insert
//   -> amplitude fix-up (returns at once unless the insert saw duplicated keys).
// G spreads the keys that share an alpha half over 2^G lines (selected by hash bits of the beta half) when sample
// sets concentrate on few alpha strings; G = 0 when at most 5 % of the keys sit in lines holding more than 128 keys.
#include <algorithm>

#include "common.cuh"

namespace anqs {

// The three GF(2)-linear hashes of a key, per thread: a loop over the set bits of each half (N_alpha / N_beta of them).
// These loops are ~70 % of the insert kernel's instructions.  Measured alternative: byte tables in shared memory (four loads per
// hash, 12 KB filled by every block first) made the build SLOWER (1.237 against 1.207 ms for 8.4M keys, 0.140 against 0.117 ms for
// 1M): the kernel waits on its atomics, not on the issue slots, and the table fill delays every block's first key
// (profiles/r2_table_build.txt).
__device__ __forceinline__ void key_hashes(uint64_t key, uint32_t &hl, uint32_t &hp) {
    const uint32_t ka = (uint32_t)key, kb = (uint32_t)(key >> 32);
    hl = lin_dev(LIN_LINE, ka);
    hp = (lin_dev(LIN_POSA, ka) & POSA_MASK) ^ (lin_dev(LIN_POSB, kb) & POSB_MASK);
}

__global__ void __launch_bounds__(256)
filter_count_kernel(const int64_t *__restrict__ keys, int64_t n, uint32_t *counts, uint32_t linemask) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        uint32_t hl, hp;
        key_hashes(deinterleave((uint64_t)keys[j]), hl, hp);
        atomicAdd(counts + (hl & linemask), 1u);
    }
}

// overloaded[g] = number of keys in alpha-lines that hold more than 128 << g keys; G = first g whose overloaded share is
// <= 5 %.  Two launches: a grid-wide sum into 64-bit accumulators kept in the spare bytes of the header, then one thread
// that picks G (a single block used to scan all the line counters: 0.13 ms at 8M keys).
__global__ void __launch_bounds__(256)
filter_overload_kernel(const uint32_t *__restrict__ counts, uint32_t nlines, unsigned long long *acc) {
    __shared__ unsigned long long sacc[FILTER_MAX_SPREAD_BITS + 1];
    if (threadIdx.x <= FILTER_MAX_SPREAD_BITS) sacc[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long local[FILTER_MAX_SPREAD_BITS + 1] = {0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nlines; i += gridDim.x * blockDim.x) {
        const uint32_t c = counts[i];
#pragma unroll
        for (int g = 0; g <= FILTER_MAX_SPREAD_BITS; ++g)
            if (c > (128u << g)) local[g] += c;
    }
#pragma unroll
    for (int g = 0; g <= FILTER_MAX_SPREAD_BITS; ++g)
        if (local[g]) atomicAdd(&sacc[g], local[g]);
    __syncthreads();
    if (threadIdx.x <= FILTER_MAX_SPREAD_BITS && sacc[threadIdx.x]) atomicAdd(&acc[threadIdx.x], sacc[threadIdx.x]);
}

__global__ void filter_pick_spread_kernel(unsigned long long *__restrict__ acc, uint32_t nlines, uint32_t n_keys,
                                          FilterHeader *hdr, int forced_spread) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int G = FILTER_MAX_SPREAD_BITS;
    for (int g = FILTER_MAX_SPREAD_BITS; g >= 0; --g) {
        const unsigned long long a = forced_spread < 0 ? acc[g] : 0ull;
        hdr->overloaded[g] = (uint32_t)a;
        if (a * 20ull <= (unsigned long long)n_keys) G = g;
    }
    if (forced_spread >= 0) G = forced_spread;
    while (G > 0 && (1u << G) > nlines) --G;
    hdr->spread_bits = (uint32_t)G;
    hdr->gmask = (1u << G) - 1u;
    hdr->n_keys = n_keys;
    acc[0] = 0ull;   // from here on the first accumulator is the insert kernels' "duplicate keys seen" flag
}

// Insert kernel: keys into the table and the filter.  An insert is a chain of dependent round trips to the L2 / DRAM (key ->
// compare-and-swap on the home slot -> maximum on the position -> amplitude stores) and the kernel is latency-bound (ncu, 8.4M
// keys: 36-43 warps stalled on the long scoreboard per issue, DRAM at 30 %, L2 at 20 %; profiles/r2_table_build.txt).  Tried
// and measured slower: four keys per thread side by side (57 registers halve the resident warps and the probing loops of the
// four keys run one after the other: 1.35 against 1.0 ms), keys grouped by home slot into L2-sized partitions first (same DRAM
// traffic - every sector still comes from DRAM once after the memset - plus 0.2 ms of partitioning).
// Duplicates: the largest position wins, which is what a sequential scatter_ leaves behind, and the amplitude has to follow the
// position.  Every thread that raises idx stores its amplitude; when two positions of one key race, the stores may land in
// either order - so a thread that finds the slot already claimed (old position >= 0) raises *dup_flag, and
// hash_fix_amplitudes_kernel, launched behind this kernel, then rewrites every slot's amplitude from its final position.
// No fence and no re-read inside the insert (__threadfence() is a gpu-scope MEMBAR plus an L1 invalidation per key), and
// batches without duplicates - the normal case: the sampler returns unique configurations - never run the second pass.
__global__ void __launch_bounds__(256)
hash_build_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                  uint32_t *filter_words, const FilterHeader *hdr, unsigned long long *dup_flag, uint32_t capmask, uint32_t linemask) {
    const uint32_t gmask = hdr->gmask;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool dup = false;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t key = deinterleave((uint64_t)keys[j]);
        uint32_t hl, hp;
        key_hashes(key, hl, hp);
        const uint32_t line = (hl ^ ((hp >> 15) & gmask)) & linemask;
        atomicOr(filter_words + (size_t)line * 32 + ((hp >> 5) & 31u), (1u << (hp & 31u)) | (1u << ((hp >> 10) & 31u)));
        HashSlot *sl;
        if (key == EMPTY_KEY) {   // the all-ones key is the EMPTY sentinel: it lives in a dedicated slot behind the table
            sl = slots + (size_t)capmask + 1;
        } else {
            uint32_t h = hash_key((uint32_t)key, (uint32_t)(key >> 32)) & capmask;
            for (;;) {
                unsigned long long prev = atomicCAS((unsigned long long *)&slots[h].key, (unsigned long long)EMPTY_KEY,
                                                    (unsigned long long)key);
                if (prev == EMPTY_KEY || prev == key) break;
                h = (h + 1) & capmask;   // occupied by another key: linear probing
            }
            sl = slots + h;
        }
        const long long old = atomicMax(&sl->idx, (long long)j);
        if (old < (long long)j && amps) {
            const double2 a = amps[j];
            sl->re = a.x;
            sl->im = a.y;
        }
        dup |= old >= 0;
    }
    if (dup) *dup_flag = 1ull;
}

// Second pass, only when the insert saw duplicates: position j rewrites the amplitude of its key's slot if it is the winner.
__global__ void __launch_bounds__(256)
hash_fix_amplitudes_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                           const unsigned long long *__restrict__ dup_flag, uint32_t capmask) {
    if (*dup_flag == 0ull) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t key = deinterleave((uint64_t)keys[j]);
        HashSlot *sl;
        if (key == EMPTY_KEY) {
            sl = slots + (size_t)capmask + 1;
        } else {
            uint32_t h = hash_key((uint32_t)key, (uint32_t)(key >> 32)) & capmask;
            while (slots[h].key != key) h = (h + 1) & capmask;   // present: the insert kernel has finished
            sl = slots + h;
        }
        if (sl->idx == (long long)j) {
            const double2 a = amps[j];
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
        double re, im;
        long long j = hash_lookup(hv, q, re, im);
        if (ptr) ptr[i] = j;
        if (mask) mask[i] = j >= 0 ? 1 : 0;
    }
}

}  // namespace anqs

using namespace anqs;

static int build_table(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                       int forced_spread, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative key count");
    ANQS_REQUIRE(d_table, "null table buffer");
    ANQS_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 1024");
    ANQS_REQUIRE(capacity >= 2 * n, "capacity must be at least 2n (use anqs_hash_capacity)");
    ANQS_REQUIRE(capacity <= ((int64_t)1 << 29), "capacity above 2^29 slots is not supported");
    ANQS_REQUIRE(((uintptr_t)d_table & 127) == 0, "table buffer must be 128-byte aligned");
    ANQS_REQUIRE(forced_spread <= FILTER_MAX_SPREAD_BITS, "spread_bits must be at most 6");
    cudaStream_t s = (cudaStream_t)stream;
    HashView hv = make_hash_view(d_table, capacity);
    HashSlot *slots = (HashSlot *)d_table;
    FilterHeader *hdr = (FilterHeader *)hv.header;
    uint32_t *filter_words = (uint32_t *)hv.filter;
    const uint32_t nlines = hv.linemask + 1;
    // 0xFF over the slots: key = EMPTY, idx = -1; zero header and filter
    ANQS_CUDA(cudaMemsetAsync(slots, 0xFF, (size_t)(capacity + 1) * sizeof(HashSlot), s));
    ANQS_CUDA(cudaMemsetAsync(hdr, 0, 96, s));
    ANQS_CUDA(cudaMemsetAsync(filter_words, 0, (size_t)FILTER_BYTES_PER_SLOT * capacity, s));
    if (n == 0) return 0;
    ANQS_REQUIRE(d_keys, "null key array");
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
    if (forced_spread < 0) {
        // per-line key counts, kept in the (still empty) filter region: nlines * 4 bytes <= the filter's size
        filter_count_kernel<<<grid, 256, 0, s>>>(d_keys, n, filter_words, hv.linemask);
        ANQS_LAUNCH_CHECK();
    }
    // 64-bit accumulators in the spare bytes of the 96-byte header (zeroed by the memset above)
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(hdr) + 40);
    if (forced_spread < 0) {
        filter_overload_kernel<<<std::min<uint32_t>((nlines + 255) / 256, 1024u), 256, 0, s>>>(filter_words, nlines, acc);
        ANQS_LAUNCH_CHECK();
    }
    filter_pick_spread_kernel<<<1, 32, 0, s>>>(acc, nlines, (uint32_t)n, hdr, forced_spread);
    ANQS_LAUNCH_CHECK();
    if (forced_spread < 0) ANQS_CUDA(cudaMemsetAsync(filter_words, 0, (size_t)nlines * sizeof(uint32_t), s));
    hash_build_kernel<<<grid, 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, slots, filter_words, hdr, acc, hv.capmask, hv.linemask);
    ANQS_LAUNCH_CHECK();
    if (d_amps != nullptr) {
        hash_fix_amplitudes_kernel<<<std::min(grid, sm_count_of_current_device() * 8), 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, slots, acc, hv.capmask);
        ANQS_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" {

int64_t anqs_hash_capacity(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

size_t anqs_hash_bytes(int64_t capacity) {
    return (size_t)capacity * sizeof(HashSlot) + 128 + FILTER_ALIGN + (size_t)FILTER_BYTES_PER_SLOT * capacity;
}

int anqs_hash_build(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                    void *stream) {
    return build_table(d_keys, d_amps, n, d_table, capacity, -1, stream);
}

int anqs_hash_build_spread(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                           int spread_bits, void *stream) {
    return build_table(d_keys, d_amps, n, d_table, capacity, spread_bits, stream);
}

int anqs_hash_filter_info(const void *d_table, int64_t capacity, int *spread_bits, int64_t *overloaded_keys, void *stream) {
    ANQS_REQUIRE(d_table, "null table buffer");
    ANQS_REQUIRE(capacity >= 1024 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two >= 1024");
    HashView hv = make_hash_view(d_table, capacity);
    FilterHeader h;
    ANQS_CUDA(cudaMemcpyAsync(&h, hv.header, sizeof(FilterHeader), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ANQS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (spread_bits) *spread_bits = (int)h.spread_bits;
    if (overloaded_keys)
        for (int g = 0; g <= FILTER_MAX_SPREAD_BITS; ++g) overloaded_keys[g] = (int64_t)h.overloaded[g];
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
