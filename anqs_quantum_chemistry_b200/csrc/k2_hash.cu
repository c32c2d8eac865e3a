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
// Build = 5 stream-ordered launches, no host synchronisation:
//   memset -> count keys per line (in the filter region itself) -> pick the spread G -> memset filter -> insert.
// G spreads the keys that share an alpha half over 2^G lines (selected by hash bits of the beta half) when sample
// sets concentrate on few alpha strings; G = 0 when at most 5 % of the keys sit in lines holding more than 128 keys.
#include <algorithm>

#include "common.cuh"

namespace anqs {

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

__global__ void filter_pick_spread_kernel(const unsigned long long *__restrict__ acc, uint32_t nlines, uint32_t n_keys,
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
}

// One key into the table and the filter.  `first` = the amplitude of position j when the caller already holds it (a record
// of the partitioned build), else it is read from amps[j].
__device__ __forceinline__ void insert_key(uint64_t key, long long j, const double2 *first, const double2 *__restrict__ amps,
                                           HashSlot *slots, uint32_t *filter_words, uint32_t gmask, uint32_t capmask,
                                           uint32_t linemask) {
    uint32_t hl, hp;
    key_hashes(key, hl, hp);
    const uint32_t line = (hl ^ ((hp >> 15) & gmask)) & linemask;
    atomicOr(filter_words + (size_t)line * 32 + ((hp >> 5) & 31u), (1u << (hp & 31u)) | (1u << ((hp >> 10) & 31u)));
    HashSlot *sl;
    if (key == EMPTY_KEY) {
        sl = slots + (size_t)capmask + 1;
    } else {
        uint32_t h = hash_key((uint32_t)key, (uint32_t)(key >> 32)) & capmask;
        for (;;) {
            unsigned long long prev = atomicCAS((unsigned long long *)&slots[h].key, (unsigned long long)EMPTY_KEY,
                                                (unsigned long long)key);
            if (prev == EMPTY_KEY || prev == key) break;
            h = (h + 1) & capmask;
        }
        sl = slots + h;
    }
    // duplicates: the largest position wins, which is what a sequential scatter_ leaves behind.  The amplitude has to follow
    // the position: a thread that raised idx writes its amplitude, then re-reads idx and, if a larger position has arrived
    // meanwhile, writes THAT position's amplitude - so whichever store lands last carries the amplitude of the final idx.
    long long old = atomicMax(&sl->idx, j);
    if (old < j && amps) {
        long long cur = j;
        double2 a = first ? *first : amps[cur];
        for (;;) {
            sl->re = a.x;
            sl->im = a.y;
            __threadfence();
            const long long now = *reinterpret_cast<volatile long long *>(&sl->idx);
            if (now == cur) break;
            cur = now;
            a = amps[cur];
        }
    }
}

__global__ void __launch_bounds__(256)
hash_build_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                  uint32_t *filter_words, const FilterHeader *hdr, uint32_t capmask, uint32_t linemask) {
    const uint32_t gmask = hdr->gmask;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        insert_key(deinterleave((uint64_t)keys[j]), (long long)j, nullptr, amps, slots, filter_words, gmask, capmask, linemask);
}

// ---- partitioned build for tables that do not fit the L2 cache -------------------------------------------------------------
// A direct build of an 8.4M-key table (537 MB of slots) sends every key to a random 32-byte sector of DRAM and back: 1.12 ms on
// B200, 7.5 G keys/s - the rate of random DRAM accesses, not a bandwidth (profiles/r2_table_build.txt).  Here the keys are
// first grouped by the leading bits of their home slot into partitions of PART_BYTES of slots (three streaming passes:
// histogram, scan, scatter of 32-byte records {key, position, amplitude}), and the insert kernel walks the records in that
// order: at any moment its threads work on one or two partitions, which the L2 holds, and the slots reach DRAM as whole
// lines when they are evicted.  The order of the records inside a partition is arbitrary; the result does not depend on it
// (largest position wins, amplitude follows).
constexpr int PART_CHUNK = 8;                     // keys per thread of the histogram / scatter kernels
constexpr int PART_MAX = 1024;
constexpr size_t PART_BYTES = (size_t)16 << 20;   // slots per partition, in bytes

struct __align__(16) KeyRecord {
    uint64_t key;   // de-interleaved
    long long j;
    double re, im;
};

__device__ __forceinline__ uint32_t home_partition(uint64_t key, uint32_t capmask, int part_shift) {
    return key == EMPTY_KEY ? 0u : (hash_key((uint32_t)key, (uint32_t)(key >> 32)) & capmask) >> part_shift;
}

__global__ void __launch_bounds__(256)
part_count_kernel(const int64_t *__restrict__ keys, int64_t n, uint32_t capmask, int part_shift, int parts,
                  unsigned long long *__restrict__ counts) {
    __shared__ uint32_t hist[PART_MAX];
    for (int p = threadIdx.x; p < parts; p += blockDim.x) hist[p] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * PART_CHUNK);
#pragma unroll
    for (int c = 0; c < PART_CHUNK; ++c) {
        const int64_t j = base + c * 256 + threadIdx.x;
        if (j < n) atomicAdd(&hist[home_partition(deinterleave((uint64_t)keys[j]), capmask, part_shift)], 1u);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < parts; p += blockDim.x)
        if (hist[p]) atomicAdd(&counts[p], (unsigned long long)hist[p]);
}

// exclusive scan of the partition sizes into cursors (one block; at most 1 024 partitions)
__global__ void __launch_bounds__(PART_MAX)
part_scan_kernel(const unsigned long long *__restrict__ counts, int parts, unsigned long long *__restrict__ cursors) {
    __shared__ unsigned long long sh[PART_MAX];
    const int t = threadIdx.x;
    sh[t] = t < parts ? counts[t] : 0ull;
    __syncthreads();
    for (int d = 1; d < PART_MAX; d <<= 1) {
        const unsigned long long v = t >= d ? sh[t - d] : 0ull;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    if (t < parts) cursors[t] = sh[t] - counts[t];
}

__global__ void __launch_bounds__(256)
part_scatter_kernel(const int64_t *__restrict__ keys, const double2 *__restrict__ amps, int64_t n, uint32_t capmask,
                    int part_shift, int parts, unsigned long long *__restrict__ cursors, KeyRecord *__restrict__ records) {
    __shared__ uint32_t hist[PART_MAX];
    __shared__ unsigned long long start[PART_MAX];
    for (int p = threadIdx.x; p < parts; p += blockDim.x) hist[p] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * PART_CHUNK);
    uint64_t key[PART_CHUNK];
    uint32_t part[PART_CHUNK], rank[PART_CHUNK];
#pragma unroll
    for (int c = 0; c < PART_CHUNK; ++c) {
        const int64_t j = base + c * 256 + threadIdx.x;
        if (j < n) {
            key[c] = deinterleave((uint64_t)keys[j]);
            part[c] = home_partition(key[c], capmask, part_shift);
            rank[c] = atomicAdd(&hist[part[c]], 1u);
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < parts; p += blockDim.x)
        if (hist[p]) start[p] = atomicAdd(&cursors[p], (unsigned long long)hist[p]);
    __syncthreads();
#pragma unroll
    for (int c = 0; c < PART_CHUNK; ++c) {
        const int64_t j = base + c * 256 + threadIdx.x;
        if (j < n) {
            KeyRecord r;
            r.key = key[c];
            r.j = (long long)j;
            const double2 a = amps ? amps[j] : make_double2(0.0, 0.0);
            r.re = a.x;
            r.im = a.y;
            records[start[part[c]] + rank[c]] = r;
        }
    }
}

__global__ void __launch_bounds__(256)
hash_build_records_kernel(const KeyRecord *__restrict__ records, const double2 *__restrict__ amps, int64_t n, HashSlot *slots,
                          uint32_t *filter_words, const FilterHeader *hdr, uint32_t capmask, uint32_t linemask) {
    const uint32_t gmask = hdr->gmask;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(records + i));
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(records + i) + 1);
        const uint64_t key = ((uint64_t)lo.y << 32) | lo.x;
        const long long j = (long long)(((uint64_t)lo.w << 32) | lo.z);
        const double2 a = make_double2(__hiloint2double((int)hi.y, (int)hi.x), __hiloint2double((int)hi.w, (int)hi.z));
        insert_key(key, j, &a, amps, slots, filter_words, gmask, capmask, linemask);
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

// number of partitions of the partitioned build, 0 = the table is small enough for the direct build
static int partition_count(int64_t capacity) {
    const size_t slot_bytes = (size_t)capacity * sizeof(HashSlot);
    if (slot_bytes < ((size_t)96 << 20)) return 0;
    return (int)std::min<size_t>(slot_bytes / PART_BYTES, (size_t)PART_MAX);
}
static size_t partition_workspace(int64_t n, int64_t capacity) {
    return partition_count(capacity) == 0 ? 0 : (size_t)n * sizeof(KeyRecord) + 2 * PART_MAX * sizeof(unsigned long long);
}

static int build_table(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                       int forced_spread, void *stream, void *d_work = nullptr, size_t work_bytes = 0) {
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
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count_of_current_device() * 16);
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
    const int parts = partition_count(capacity);
    if (parts > 0 && d_work != nullptr && work_bytes >= partition_workspace(n, capacity)) {
        ANQS_REQUIRE(((uintptr_t)d_work & 15) == 0, "workspace must be 16-byte aligned");
        KeyRecord *records = (KeyRecord *)d_work;
        unsigned long long *counts = (unsigned long long *)(records + n), *cursors = counts + PART_MAX;
        int log2cap = 0;
        while (((int64_t)1 << log2cap) < capacity) ++log2cap;
        int log2parts = 0;
        while ((1 << log2parts) < parts) ++log2parts;
        const int part_shift = log2cap - log2parts;
        ANQS_CUDA(cudaMemsetAsync(counts, 0, 2 * PART_MAX * sizeof(unsigned long long), s));
        const int pgrid = (int)((n + 256 * PART_CHUNK - 1) / (256 * PART_CHUNK));
        part_count_kernel<<<pgrid, 256, 0, s>>>(d_keys, n, hv.capmask, part_shift, parts, counts);
        ANQS_LAUNCH_CHECK();
        part_scan_kernel<<<1, PART_MAX, 0, s>>>(counts, parts, cursors);
        ANQS_LAUNCH_CHECK();
        part_scatter_kernel<<<pgrid, 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, hv.capmask, part_shift, parts, cursors, records);
        ANQS_LAUNCH_CHECK();
        hash_build_records_kernel<<<grid, 256, 0, s>>>(records, (const double2 *)d_amps, n, slots, filter_words, hdr, hv.capmask,
                                                       hv.linemask);
        ANQS_LAUNCH_CHECK();
        return 0;
    }
    hash_build_kernel<<<grid, 256, 0, s>>>(d_keys, (const double2 *)d_amps, n, slots, filter_words, hdr, hv.capmask, hv.linemask);
    ANQS_LAUNCH_CHECK();
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

size_t anqs_hash_build_workspace(int64_t n, int64_t capacity) {
    return n <= 0 ? 0 : partition_workspace(n, capacity);
}

int anqs_hash_build_ws(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity, int spread_bits,
                       void *d_work, size_t work_bytes, void *stream) {
    return build_table(d_keys, d_amps, n, d_table, capacity, spread_bits, stream, d_work, work_bytes);
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
