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

// Product-layout records (see Tables below)
struct ProdTile {
    uint32_t blob_off;     // byte offset of the tile in prod_blob (multiple of 128)
    uint32_t blob_bytes;   // multiple of 16
    uint32_t n_multi, n_single, n_members;
    uint32_t row_base;     // global index of the tile's first row (into prod_row_u)
    uint32_t member_base;  // global index of the tile's first member (into prod_mem_u)
    uint32_t pad;
};
struct RowRec {            // 16 bytes
    uint32_t pa;           // alpha part of the masks of this row
    uint32_t hline;        // lin(LIN_LINE, pa)
    uint32_t a;            // multi: first member (tile-local index) | singleton: beta part mb
    uint32_t b;            // multi: member count                   | singleton: member hash
};
struct MemRec {            // 8 bytes
    uint32_t mb;           // beta part
    uint32_t hash;         // lin(LIN_POSA, pa) ^ lin(LIN_POSB, mb)
};
// Enumeration tile (k1_enum.cu): a contiguous range of XY masks [u0, u0 + n_masks) with everything the matrix elements of
// its PATTERN groups need, sized to live in shared memory.  Blob layout (one bulk copy), every section 16-byte aligned:
//   [rec:  n_masks x 16 B = {xy, mult, off}]   one LDS.128 per connection
//   [pattern tables: n_tab x 8 B (re)][n_tab x 8 B (im) when weights are complex]
//   [occupation blocks of the generic groups]
//   [generic-mask bitmap: n_words x 4 B, bit b of word j = mask 32 j + b is generic]
// A YZ group is of PATTERN type when all its terms differ only on the XY positions of its mask (true for every
// four-position group of a Jordan-Wigner molecular Hamiltonian): then
//   H_{x,x'} = (-1)^parity(x & zbase) * G[slot],   slot = (((v.hi * ENUM_FOLD + v.lo) * mult) >> 29),  v = x' & xy,
// with G summed at table-build time and `mult` an odd 32-bit multiplier searched per mask so that the occupation patterns a
// sample of the (N_alpha, N_beta) sector can show on the mask's positions (exactly half of the alpha positions and half of
// the beta positions occupied: 2, 4 or 6 patterns) land on distinct slots of a table of 2, 4 or 8 doubles - five integer
// instructions instead of three variable bit extractions.  `off` is the byte offset of G inside the tile.  zbase itself is
// not stored: for a Jordan-Wigner string parity(x & zbase) = parity(S & xy) up to a pattern-dependent sign folded into G,
// S = the exclusive prefix parity of the sample, computed once per (sample, tile) unit (analyse_group in abi_core.cu).
// Generic groups (the diagonal, one-body excitations dressed with number operators; ~1 % of the connections) and pattern
// groups whose zbase is not of that form have mult = 0; `off` then points at an OCCUPATION BLOCK in the tile,
// [mult u32][kind | nbits << 8 u32][zbase u64][tables]:
//   kind 2: the Z parts of the terms outside the mask differ from zbase by at most one position r:
//           H = sign * (A[slot] + sum over the occupied positions r of x' of D[r][slot]),  tables A[2^nbits], D[n][2^nbits]
//   kind 3: xy = 0 and at most two Z positions per term:  H = K + sum_i n_i (a_i + sum_{j<i} n_j b_ij),  tables K, a[n], b[n][n]
//   kind 4: a pattern group with an explicit zbase:  H = sign * G[slot]
// (complex weights: the imaginary tables follow the real ones), or off = 0 when the group is none of these: its term records
// are then summed from the global arrays.  Generic connections are deferred by the emit kernel (a per-warp list in global
// memory) and evaluated 32 at a time, one connection per lane.  Samples outside the (N_alpha, N_beta) sector, for which no
// table applies, take a separate slow path through the global term arrays.
struct EnumTile {
    uint32_t blob_off;     // byte offset in enum_blob (multiple of 128)
    uint32_t blob_bytes;   // multiple of 16
    uint32_t u0, n_masks;  // u0 and n_masks are multiples of 32
    uint32_t n_tab, n_generic;
    uint32_t desc_off, tab_off;  // byte offsets inside the tile
    uint32_t word0, n_words;     // the tile's slice of a bitmap row
    uint32_t gen_off, pad1;      // gen_off: byte offset of the per-word bitmaps of the generic masks (n_words x 4 B)
};
constexpr uint32_t ENUM_FOLD = 0x9E3779B1u;  // folds the two words of x' & xy into one before the per-mask multiplier
// shared-memory budget of enum_emit_kernel: per-warp queues of tile-local mask indices, one resident tile
constexpr int ENUM_EMIT_WARPS = 32;
constexpr int ENUM_STEP_WORDS = 32;                        // bitmap words expanded per step
constexpr int ENUM_QCAP = ENUM_STEP_WORDS * 32 + 32;       // queued indices per warp
constexpr int ENUM_QUEUE_BYTES = ENUM_EMIT_WARPS * ENUM_QCAP * 2;
constexpr int ENUM_DEFER_CAP = 64;                         // deferred connections per warp (global workspace, 16 B each)
constexpr uint32_t ENUM_TILE_MAX = 227 * 1024 - ENUM_QUEUE_BYTES - 1024;  // bytes one enumeration tile may take
constexpr int ENUM_PATTERN_MAX_TERMS = 64;     // longer groups stay generic (their tables would not save anything)
// Device-resident Hamiltonian tables (reference tensors PO:103-115, re-laid-out for the kernels).
struct Tables {
    int qubit_num;
    int weights_real;
    int device;
    int64_t U, T;
    int64_t U_pad;           // U rounded up to a multiple of 1024 (32 bitmap words)
    int64_t row_words;       // U_pad / 32
    int max_group;           // largest YZ group
    int max_xy_weight;       // largest popcount of an XY mask (the pair-join kernel's pruning radius)
    uint64_t *xy;            // [U_pad]  unique XY masks, ascending (signed order, as torch.unique gives)
    uint2 *mab;              // [U_pad]  de-interleaved masks: .x = even (alpha) bits, .y = odd (beta) bits
    int2 *grp;               // [U_pad]  (start, num) of the YZ group of each XY mask
    uint64_t *yz_d;          // [T]  YZ masks, de-interleaved: low word = even (alpha) bits, high word = odd bits
    double *w_re;            // [T]
    double *w_im;            // [T] (NULL when weights_real)
    ulonglong2 *term_real;   // [T] packed {yz_d, bits(w_re)} records for one 16-byte load (weights_real only)

    // ---- product layout (k1_fused.cu): masks factorised into (alpha part, beta part) ---------------------
    // A "row" is one distinct alpha part pa together with the beta parts that occur with it.  The electron-count
    // filter factorises, popc(xa ^ pa) == N_alpha and popc(xb ^ mb) == N_beta, so a row whose alpha part fails
    // is skipped as a whole.  Rows are packed into tiles (one bulk copy each):
    //   [RowRec x (n_multi + n_single)] [MemRec x n_members]
    // multi rows list their members in the tile's MemRec array; rows with <= 2 members are expanded to
    // "singleton" records that carry their one member inline.
    int n_tiles;
    int tile_bytes_max;
    ProdTile *prod_tiles;     // [n_tiles] directory (device)
    uint8_t *prod_blob;       // tile blobs, each 128-byte aligned (device)
    uint32_t *prod_row_u;     // [total rows]    mask index u of a singleton row (device)
    uint32_t *prod_mem_u;     // [total members] mask index u of a member (device)

    // ---- enumeration layouts (k1_enum.cu) ------------------------------------------------------------------
    uint8_t *prod_blob_u;     // same tiles as prod_blob with the mask index u in place of the hashes
                              // (MemRec::hash and the singleton RowRec::b)
    uint8_t *prod_blob_bs;    // same tiles with slice positions (part_positions in abi_core.cu) in place of RowRec::pa,
                              // the singleton RowRec::a and MemRec::mb: the records of the bit-sliced fused kernel
    int prod_bs_ok;           // 0: a spin part of some mask has an even weight > 4 (the bit-sliced kernels do not apply)
    // Bit-sliced filter: per mask the bit positions (in x) of its alpha part (4 bytes) and beta part (4 bytes), padded with
    // the constant slices 64 (all zeros) / 65 (all ones) so that "exactly two of the four slices set" is the electron-count
    // test of every part of weight 0, 2 or 4 (weight 0: {0,0,1,1}; weight 2: {p,q,0,1}; odd weights: {0,0,0,0} = never).
    uint2 *bs_pos;            // [U_pad]
    int bs_ok;                // 0: some part has an even weight > 4, the bit-sliced filter does not apply
    int n_enum_tiles;         // 0: the tiled enumeration is unavailable for this table
    int enum_tile_bytes_max;
    EnumTile *enum_tiles;     // [n_enum_tiles] directory (device)
    uint8_t *enum_blob;       // tile blobs, each 128-byte aligned (device)
    Tables *dev_copy;         // this struct in device memory: what rarely-taken non-inlined device functions read the table pointers from
};

// ---- sampled-set lookup table (kernel family 2) ------------------------------------------------------
// Memory layout of the caller-allocated buffer (128-byte aligned):
//   [capacity slots of 32 bytes][dedicated slot for the all-ones key, 32 bytes][header, 96 bytes][pad to an 8 KB boundary][filter]
// The filter is a line-blocked presence filter of 2*capacity bytes (32..64 bits per key, TWO bits of one word per key):
// the 128-byte line is chosen by the ALPHA half of the key (lin LIN_LINE), optionally spread over 2^G lines by
// G hash bits of the beta half, and the bit inside the line by both halves (LIN_POSA ^ LIN_POSB).  All the hashes
// are GF(2)-linear, so the fused kernel gets the filter address of x' = x ^ mask with one XOR per candidate, and
// every candidate that shares the sample and the alpha part of the mask (a whole row of the product layout) tests a
// bit of the SAME line: one L1 wavefront per warp step instead of one per candidate.  G is chosen at build time
// from the occupancy of the lines (k2_hash.cu) and stored in the header.  Keys are stored DE-INTERLEAVED.
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFULL;  // de-interleaving maps all-ones to all-ones
constexpr int FILTER_MAX_SPREAD_BITS = 6;
constexpr int FILTER_BYTES_PER_SLOT = 2;  // filter size = 2 * capacity bytes: 32..64 bits per key, TWO bits of one word per key (~0.3 % false positives)
constexpr uint32_t POSA_MASK = 0x7FFFu, POSB_MASK = 0x1FFFFFu;  // widths of the two position hashes
constexpr size_t FILTER_ALIGN = 8192;  // > every member part of a probe offset: 6 spread bits << 7 | 5 word bits << 2
struct __align__(32) HashSlot {
    uint64_t key;    // de-interleaved configuration
    long long idx;   // position in the key array (-1 = empty)
    double re, im;   // amplitude psi(key)
};
struct FilterHeader {  // lives in the 96 bytes behind the dedicated slot
    uint32_t spread_bits;   // G
    uint32_t gmask;         // (1 << G) - 1
    uint32_t overloaded[FILTER_MAX_SPREAD_BITS + 1];  // keys in lines holding more than 128 << g keys, per candidate g
    uint32_t n_keys;
    // bytes 40..95 of the header region: seven 64-bit accumulators of the overload sums while the table is being built
};
struct HashView {
    const HashSlot *slots;
    const FilterHeader *header;
    const uint8_t *filter;   // nlines * 128 bytes
    uint32_t capmask;        // capacity - 1
    uint32_t linemask;       // nlines - 1, nlines = capacity * FILTER_BYTES_PER_SLOT / 128
};
struct Tables;
// k1_fused_bs.cu: launches the bit-sliced fused local-energy kernel; 1 = launched, 0 = does not apply, < 0 = CUDA error
int fused_bs_try_launch(const Tables *t, HashView hv, const int64_t *d_samples, const double *d_amps, int64_t row_start,
                        int64_t row_len, int alpha_num, int beta_num, double *d_eloc, int variant, cudaStream_t s);
inline HashView make_hash_view(const void *d_table, int64_t capacity) {
    HashView hv;
    hv.slots = (const HashSlot *)d_table;
    hv.header = (const FilterHeader *)(hv.slots + capacity + 1);
    // the filter starts at the first FILTER_ALIGN boundary behind the header: the fused kernels form probe addresses as
    // (filter + sample part) ^ (member part < FILTER_ALIGN), which equals filter + (sample part ^ member part) only then
    hv.filter = (const uint8_t *)(((uintptr_t)(hv.slots + capacity) + 128 + (FILTER_ALIGN - 1)) & ~(uintptr_t)(FILTER_ALIGN - 1));
    hv.capmask = (uint32_t)(capacity - 1);
    hv.linemask = (uint32_t)(capacity * FILTER_BYTES_PER_SLOT / 128 - 1);
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
// ---- GF(2)-linear hashes of the two halves of a de-interleaved configuration ------------------------------
// lin(f, v) = XOR of LIN_C[f][i] over the set bits i of v, so lin(f, x ^ m) = lin(f, x) ^ lin(f, m): the hash of a
// connected configuration x' = x ^ mask is the hash of the sample XOR a per-mask constant precomputed at table
// build time.  Columns are fixed random constants (full-rank matrices), identical on host and device.
#define ANQS_LIN_TABLE                                                                                                    \
    {                                                                                                                     \
        /* LIN_LINE: alpha half -> 32-bit line hash */                                                                    \
        {0xb2285d19u, 0xc35cafefu, 0xb18c34eeu, 0x2c91baccu, 0x2ede2defu, 0x06f094b1u, 0xe5fb86d2u, 0xd176b960u,          \
         0x810729c9u, 0x22bb38deu, 0xfa9dbac4u, 0x11ab6a6du, 0x81d0ff89u, 0x1e7b2ca5u, 0x92eea3a6u, 0x24949e26u,          \
         0x90f3f271u, 0x68f545b0u, 0x2e32da50u, 0xd9779982u, 0x0712e2ccu, 0x7ca6fa6eu, 0x09e2c4a5u, 0xd73e2794u,          \
         0x61f91774u, 0x3f9b14f2u, 0x9dd55901u, 0x05ad50e5u, 0x9400cb1cu, 0xb4e13945u, 0x3424af98u, 0x0d95e497u},         \
        /* LIN_POSA: alpha half -> bit 1 (5) | word (5) | bit 2 (5) */                                                    \
        {0x000036f5u, 0x00006d0eu, 0x0000114du, 0x0000422du, 0x00006ed5u, 0x00001a8du, 0x00004edeu, 0x00002b79u,          \
         0x000025c0u, 0x00002be8u, 0x0000195eu, 0x000056e7u, 0x00005b3eu, 0x0000158cu, 0x000016d2u, 0x0000047bu,          \
         0x00000a1bu, 0x00005108u, 0x00000eb0u, 0x00001f49u, 0x00001ab1u, 0x00002125u, 0x0000419bu, 0x0000715cu,          \
         0x000063feu, 0x000060e6u, 0x0000405cu, 0x00002280u, 0x000051c9u, 0x000047c7u, 0x000066aau, 0x00003fadu},         \
        /* LIN_POSB: beta half -> bit 1 | word | bit 2 | 6 spread bits << 15 */                                           \
        {0x0006d8eeu, 0x0001edf6u, 0x00030566u, 0x001160d2u, 0x000e4a15u, 0x00006b8bu, 0x000945f7u, 0x001f9852u,          \
         0x000350f9u, 0x0010e51cu, 0x0018e6ebu, 0x000637abu, 0x0015615eu, 0x000888e0u, 0x000631d4u, 0x0000256du,          \
         0x00133b90u, 0x0010cc32u, 0x001c63f7u, 0x00080905u, 0x00052626u, 0x0000542bu, 0x00009522u, 0x0013d69du,          \
         0x0001de04u, 0x0009fba5u, 0x0012c8acu, 0x000f5319u, 0x0013d81fu, 0x000f4908u, 0x00144f4bu, 0x0013b428u}          \
    }
constexpr int LIN_LINE = 0, LIN_POSA = 1, LIN_POSB = 2;
static const uint32_t LIN_C_HOST[3][32] = ANQS_LIN_TABLE;
inline uint32_t lin_host(int f, uint32_t v) {
    uint32_t h = 0;
    for (int i = 0; i < 32; ++i)
        if ((v >> i) & 1u) h ^= LIN_C_HOST[f][i];
    return h;
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
static __device__ const uint32_t LIN_C_DEV[3][32] = ANQS_LIN_TABLE;
// per-thread evaluation (a few set bits per key: configurations have N_alpha / N_beta electrons per half)
__device__ __forceinline__ uint32_t lin_dev(int f, uint32_t v) {
    uint32_t h = 0;
    while (v) {
        int i = __ffs(v) - 1;
        v &= v - 1;
        h ^= __ldg(&LIN_C_DEV[f][i]);
    }
    return h;
}
// warp-cooperative evaluation: lane i contributes column i, one REDUX per hash; every lane gets the result
__device__ __forceinline__ uint32_t lin_warp(int f, uint32_t v) {
    const int lane = threadIdx.x & 31;
    uint32_t c = ((v >> lane) & 1u) ? __ldg(&LIN_C_DEV[f][lane]) : 0u;
    return __reduce_xor_sync(0xffffffffu, c);
}
#endif


}  // namespace anqs
