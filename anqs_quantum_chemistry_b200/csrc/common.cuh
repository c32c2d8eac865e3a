// Internal helpers shared by the sm_100a kernels of libanqs_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/anqs_b200.h"

namespace anqs {

void set_error(const std::string &msg);
int sm_count_of_current_device();

#define ANQS_REQUIRE(cond, msg)                                                         \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            ::anqs::set_error(std::string(__func__) + ": " + (msg));                    \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define ANQS_CUDA(call)                                                                 \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::anqs::set_error(std::string(__func__) + ": " #call " failed: " +          \
                              cudaGetErrorString(e__));                                 \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

#define ANQS_LAUNCH_CHECK() ANQS_CUDA(cudaGetLastError())

// Device-resident Hamiltonian tables (reference tensors PO:103-115, re-laid-out for the kernels).
struct Tables {
    int qubit_num;
    int weights_real;
    int device;
    int64_t U, T;
    int64_t U_pad;           // U rounded up to a multiple of 1024 (32 bitmap words)
    int64_t row_words;       // U_pad / 32
    int max_group;           // largest YZ group
    uint64_t *xy;            // [U_pad]  unique XY masks, ascending (signed order, as torch.unique gives)
    uint2 *mab;              // [U_pad]  de-interleaved masks: .x = even (alpha) bits, .y = odd (beta) bits
    int2 *grp;               // [U_pad]  (start, num) of the YZ group of each XY mask
    uint64_t *yz_d;          // [T]  YZ masks, de-interleaved: low word = even (alpha) bits, high word = odd bits
    double *w_re;            // [T]
    double *w_im;            // [T] (NULL when weights_real)
    ulonglong2 *term_real;   // [T] packed {yz_d, bits(w_re)} records for one 16-byte load (weights_real only)
};

// ---- sampled-set lookup table (kernel family 2) ------------------------------------------------------
// Memory layout of the caller-allocated buffer: (capacity + 1) slots of 32 bytes, then capacity bytes of
// blocked-Bloom presence bits (capacity / 4 words of 32 bits; every key sets 3 bits of ONE word).  Keys are
// stored DE-INTERLEAVED so that the fused kernel can form the probe key (xa ^ ma, xb ^ mb) without touching
// the 64-bit masks.
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;  // de-interleaving maps all-ones to all-ones
struct __align__(32) HashSlot {
    uint64_t key;    // de-interleaved configuration
    long long idx;   // position in the key array (-1 = empty)
    double re, im;   // amplitude psi(key)
};
struct HashView {
    const HashSlot *slots;
    const uint32_t *bloom;
    uint32_t capmask;    // capacity - 1
    uint32_t wordmask;   // number of Bloom words - 1
};
inline HashView make_hash_view(const void *d_table, int64_t capacity) {
    HashView hv;
    hv.slots = (const HashSlot *)d_table;
    hv.bloom = (const uint32_t *)(hv.slots + capacity + 1);
    hv.capmask = (uint32_t)(capacity - 1);
    hv.wordmask = (uint32_t)(capacity / 4 - 1);
    return hv;
}

// Even bits of v gathered into the low 32 bits ("Morton decode").
__host__ __device__ __forceinline__ uint32_t compress_even_bits(uint64_t v) {
    v &= 0x5555555555555555ULL;
    v = (v | (v >> 1)) & 0x3333333333333333ULL;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0fULL;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffULL;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffULL;
    v = (v | (v >> 16)) & 0x00000000ffffffffULL;
    return (uint32_t)v;
}

// (even bits, odd bits) of v as (low word, high word): a bijection on 64-bit values
__host__ __device__ __forceinline__ uint64_t deinterleave(uint64_t v) {
    return (uint64_t)compress_even_bits(v) | ((uint64_t)compress_even_bits(v >> 1) << 32);
}

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }

// 32-bit hash of a de-interleaved key (murmur3-style mixing of the two halves)
__host__ __device__ __forceinline__ uint32_t hash_key(uint32_t a, uint32_t b) {
    uint32_t h = a * 0xcc9e2d51u;
    h = rotl32(h, 15) * 0x1b873593u;
    uint32_t g = b * 0x85ebca6bu;
    g = rotl32(g, 13) * 0xc2b2ae35u;
    h ^= g;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
// blocked Bloom filter: word index and 3-bit pattern, both decorrelated from the slot index (= low hash bits)
__host__ __device__ __forceinline__ uint32_t bloom_word(uint32_t h) {
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    h *= 0x297a2d39u;
    h ^= h >> 15;
    return h;
}
__host__ __device__ __forceinline__ uint32_t bloom_pattern(uint32_t h) {
    uint32_t g = h * 0x9E3779B1u;
    return (1u << (g >> 27)) | (1u << ((g >> 22) & 31u)) | (1u << ((g >> 17) & 31u));
}

#ifdef __CUDACC__
// parity of popcount(v) with a single POPC: fold the two halves first
__device__ __forceinline__ uint32_t parity64(uint64_t v) {
    return __popc((uint32_t)v ^ (uint32_t)(v >> 32)) & 1u;
}

// Looks a de-interleaved key up.  Returns the position (or -1) and, on a hit, the stored amplitude.
__device__ __forceinline__ long long hash_lookup(const HashView &hv, uint64_t key, double &re, double &im) {
    if (key == EMPTY_KEY) {  // dedicated slot after the table proper
        const HashSlot *sl = hv.slots + (size_t)hv.capmask + 1;
        re = sl->re;
        im = sl->im;
        return sl->idx;
    }
    uint32_t h = hash_key((uint32_t)key, (uint32_t)(key >> 32)) & hv.capmask;
    for (;;) {
        ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2 *>(hv.slots + h));
        if (kv.x == key) {
            double2 a = __ldg(reinterpret_cast<const double2 *>(hv.slots + h) + 1);
            re = a.x;
            im = a.y;
            return (long long)kv.y;
        }
        if (kv.x == EMPTY_KEY) return -1;
        h = (h + 1) & hv.capmask;
    }
}

// w with its sign flipped when par == 1
__device__ __forceinline__ double flip_sign(double w, uint32_t par) {
    int hi = __double2hiint(w) ^ (int)(par << 31);
    return __hiloint2double(hi, __double2loint(w));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk TMA copy (cp.async.bulk -> SASS UBLKCP) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes must be a multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif

}  // namespace anqs
