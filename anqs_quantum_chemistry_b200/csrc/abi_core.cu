// libanqs_b200.so -- error handling, device check, Hamiltonian table handle, popcount, scan.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace anqs {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int sm_count_of_current_device() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- A4 popcount: one 16-byte load + two POPC pairs + one 16-byte store per thread iteration -------
__global__ void popcount_kernel(const int64_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n2 = n >> 1;
    const longlong2 *in2 = reinterpret_cast<const longlong2 *>(in);
    longlong2 *out2 = reinterpret_cast<longlong2 *>(out);
    for (int64_t k = i; k < n2; k += stride) {
        longlong2 v = in2[k];
        v.x = __popcll((unsigned long long)v.x);
        v.y = __popcll((unsigned long long)v.y);
        out2[k] = v;
    }
    if ((n & 1) && i == 0) out[n - 1] = __popcll((unsigned long long)in[n - 1]);
}

__global__ void popcount_kernel_unaligned(const int64_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        out[k] = __popcll((unsigned long long)in[k]);
}

// ---- exclusive scan (three phases; n is at most a few 1e7 so one block scans the block sums) -------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *warp_sums) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t s = lane < (SCAN_BLOCK / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        if (lane < (SCAN_BLOCK / 32)) warp_sums[lane] = s;
    }
    __syncthreads();
    int64_t base = warp > 0 ? warp_sums[warp - 1] : 0;
    *total = warp_sums[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void scan_tile_sums(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int64_t total;
    block_exclusive_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_of_tile_sums(int64_t *tile_sums, int64_t ntiles) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t carry = 0;
    for (int64_t lo = 0; lo < ntiles; lo += SCAN_BLOCK) {
        int64_t i = lo + threadIdx.x;
        int64_t v = i < ntiles ? tile_sums[i] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, &total, warp_sums);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void scan_apply(const int64_t *__restrict__ in, int64_t n, const int64_t *__restrict__ tile_offsets,
                           int64_t *__restrict__ out) {
    __shared__ int64_t warp_sums[SCAN_BLOCK / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, &total, warp_sums) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}


// ---- product layout of the XY masks (Tables::prod_*, consumed by k1_fused.cu) --------------------------------
constexpr uint32_t PROD_TILE_MAX = 150 * 1024;  // bytes of shared memory one tile may take (the bit-sliced fused kernel keeps 75 KB of per-group state beside it)
constexpr uint32_t PROD_ROW_CHUNK = 2048;       // longer rows are split so that tiles pack evenly
constexpr uint32_t PROD_SINGLE_MAX = 2;         // rows with <= this many members become singleton records

struct HostProd {
    std::vector<ProdTile> tiles;
    std::vector<uint8_t> blob, blob_u;  // blob_u: the same records with the mask index u in place of the hashes
    std::vector<uint8_t> blob_bs;       // the same records with slice positions in place of the part words (k1_fused_bs.cu)
    bool bs_ok = true;
    std::vector<uint32_t> row_u, mem_u;
    uint32_t tile_bytes_max = 0;
};

// Slice positions of one spin part (de-interleaved, bits 0..31) for the bit-sliced electron-count test: four bytes, padded
// with the constant slices 32 (all zeros) / 33 (all ones) so that "exactly two of the four slices set" is the test for
// weights 0, 2 and 4; odd weights can never pass ({32,32,32,32}); even weights > 4 are not representable (ok = false).
static uint32_t part_positions(uint32_t bits, bool *ok) {
    const int k = __builtin_popcount(bits);
    if (k & 1) return 0x20202020u;
    if (k > 4) { *ok = false; return 0x20202020u; }
    uint32_t b[4] = {32, 32, 33, 33};
    int i = 0;
    for (int pos = 0; pos < 32; ++pos)
        if ((bits >> pos) & 1) b[i++] = (uint32_t)pos;
    if (k == 2) { b[2] = 32; b[3] = 33; }
    return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
}

static void build_product_layout(const std::vector<uint2> &mab, int64_t U, HostProd &out) {
    // masks sorted by (alpha part, spread bits of the member hash, beta part)
    struct M { uint32_t pa, mb, hash, u; };
    std::vector<M> ms((size_t)U);
    for (int64_t u = 0; u < U; ++u) {
        uint32_t pa = mab[u].x, mb = mab[u].y;
        ms[u] = {pa, mb, (lin_host(LIN_POSA, pa) & POSA_MASK) ^ (lin_host(LIN_POSB, mb) & POSB_MASK), (uint32_t)u};
    }
    std::sort(ms.begin(), ms.end(), [](const M &x, const M &y) {
        if (x.pa != y.pa) return x.pa < y.pa;
        uint32_t sx = x.hash >> 15, sy = y.hash >> 15;
        if (sx != sy) return sx < sy;
        return x.mb < y.mb;
    });
    // row chunks: [begin, end) ranges of ms with one alpha part and at most PROD_ROW_CHUNK members
    struct Chunk { size_t begin, end; };
    std::vector<Chunk> multi, single;
    for (size_t i = 0; i < ms.size();) {
        size_t j = i;
        while (j < ms.size() && ms[j].pa == ms[i].pa) ++j;
        if (j - i <= PROD_SINGLE_MAX) {
            for (size_t k = i; k < j; ++k) single.push_back({k, k + 1});
        } else {
            for (size_t k = i; k < j; k += PROD_ROW_CHUNK) multi.push_back({k, std::min(j, k + (size_t)PROD_ROW_CHUNK)});
        }
        i = j;
    }
    auto chunk_bytes = [](const Chunk &c, bool is_multi) -> size_t {
        return sizeof(RowRec) + (is_multi ? (c.end - c.begin) * sizeof(MemRec) : 0);
    };
    size_t total = 0;
    for (auto &c : multi) total += chunk_bytes(c, true);
    for (auto &c : single) total += chunk_bytes(c, false);
    const size_t n_tiles = std::max<size_t>(1, (total + PROD_TILE_MAX - 1) / PROD_TILE_MAX);
    const size_t target = (total + n_tiles - 1) / n_tiles;
    // greedy packing in order (multi rows first, then singletons); a tile closes once it reaches the target
    size_t im = 0, is = 0;
    uint32_t row_base = 0, member_base = 0;
    while (im < multi.size() || is < single.size()) {
        std::vector<Chunk> tm, ts;
        size_t bytes = 0;
        while (im < multi.size() && bytes < target && bytes + chunk_bytes(multi[im], true) <= PROD_TILE_MAX) {
            bytes += chunk_bytes(multi[im], true);
            tm.push_back(multi[im++]);
        }
        if (im == multi.size()) {
            while (is < single.size() && bytes + sizeof(RowRec) <= PROD_TILE_MAX && bytes < target) {
                bytes += sizeof(RowRec);
                ts.push_back(single[is++]);
            }
        }
        ProdTile tile{};
        tile.n_multi = (uint32_t)tm.size();
        tile.n_single = (uint32_t)ts.size();
        tile.row_base = row_base;
        tile.member_base = member_base;
        std::vector<RowRec> rows, rows_u, rows_bs;
        std::vector<MemRec> mems, mems_u, mems_bs;
        for (auto &c : tm) {
            const uint32_t pa = ms[c.begin].pa;
            rows.push_back({pa, lin_host(LIN_LINE, pa), (uint32_t)mems.size(), (uint32_t)(c.end - c.begin)});
            rows_u.push_back(rows.back());
            rows_bs.push_back({part_positions(pa, &out.bs_ok), rows.back().hline, rows.back().a, rows.back().b});
            out.row_u.push_back(0);
            for (size_t k = c.begin; k < c.end; ++k) {
                mems.push_back({ms[k].mb, ms[k].hash});
                mems_u.push_back({ms[k].mb, ms[k].u});
                mems_bs.push_back({part_positions(ms[k].mb, &out.bs_ok), ms[k].hash});
                out.mem_u.push_back(ms[k].u);
            }
        }
        for (auto &c : ts) {
            const M &m = ms[c.begin];
            rows.push_back({m.pa, lin_host(LIN_LINE, m.pa), m.mb, m.hash});
            rows_u.push_back({m.pa, 0u, m.mb, m.u});
            rows_bs.push_back({part_positions(m.pa, &out.bs_ok), lin_host(LIN_LINE, m.pa), part_positions(m.mb, &out.bs_ok), m.hash});
            out.row_u.push_back(m.u);
        }
        tile.n_members = (uint32_t)mems.size();
        size_t off = (out.blob.size() + 127) / 128 * 128;
        size_t nbytes = rows.size() * sizeof(RowRec) + mems.size() * sizeof(MemRec);
        nbytes = (nbytes + 15) / 16 * 16;
        out.blob.resize(off + nbytes, 0);
        std::memcpy(out.blob.data() + off, rows.data(), rows.size() * sizeof(RowRec));
        std::memcpy(out.blob.data() + off + rows.size() * sizeof(RowRec), mems.data(), mems.size() * sizeof(MemRec));
        out.blob_bs.resize(off + nbytes, 0);
        std::memcpy(out.blob_bs.data() + off, rows_bs.data(), rows_bs.size() * sizeof(RowRec));
        std::memcpy(out.blob_bs.data() + off + rows_bs.size() * sizeof(RowRec), mems_bs.data(), mems_bs.size() * sizeof(MemRec));
        out.blob_u.resize(off + nbytes, 0);
        std::memcpy(out.blob_u.data() + off, rows_u.data(), rows_u.size() * sizeof(RowRec));
        std::memcpy(out.blob_u.data() + off + rows_u.size() * sizeof(RowRec), mems_u.data(), mems_u.size() * sizeof(MemRec));
        tile.blob_off = (uint32_t)off;
        tile.blob_bytes = (uint32_t)nbytes;
        out.tile_bytes_max = std::max(out.tile_bytes_max, tile.blob_bytes);
        out.tiles.push_back(tile);
        row_base += (uint32_t)rows.size();
        member_base += (uint32_t)mems.size();
    }
    if (out.mem_u.empty()) out.mem_u.push_back(0);
}

// ---- enumeration tiles (Tables::enum_*, consumed by k1_enum.cu) -----------------------------------------------

struct HostEnum {
    std::vector<EnumTile> tiles;
    std::vector<uint8_t> blob;
    uint32_t tile_bytes_max = 0;
    bool ok = true;
};

// Analysis of one YZ group for the tile-resident emit kernel (see EnumTile in common.cuh).
//   kind 1 (pattern):    every term has the same Z part outside the mask; table G[slot].
//   kind 2 (one extra):  the Z parts outside the mask differ from a common zbase by at most ONE position r (one-body
//                        excitations dressed with number operators): H = sign * (A[slot] + sum_{r occupied in x'} D[r][slot]).
//   kind 3 (diagonal):   xy = 0 and every term has at most two Z positions: H = K + sum_i a_i n_i + sum_{j<i} b_ij n_i n_j.
//   kind 4 (pattern, explicit sign): a pattern group whose zbase is NOT the Jordan-Wigner string the kernel derives from the
//                        mask itself (see below): evaluated like kind 1, but on the deferred path with its own zbase.
//   kind 0:              none of these; the kernel sums the term records from global memory.
// Derived sign: with S = the exclusive prefix parity of the SAMPLE x (bit i of S = parity of the bits of x below i), the
// kernel takes the sign of a kind-1 connection as parity(S & xy) = parity(x & W), W = XOR over the positions i of the mask of
// the bits below i (the union of the ranges between the 1st and 2nd, and the 3rd and 4th position, lower ends included).
// That equals parity(x & zbase) up to bits of x ON the mask's positions whenever (zbase ^ W) lies inside the mask - true for
// every Jordan-Wigner excitation string - and those bits are a function of the occupation pattern, so they are folded into
// the table.  The mask records then carry no zbase at all: one 16-byte shared-memory load per connection instead of 24.
struct GroupInfo {
    int kind = 0, nbits = 0;
    uint32_t mult = 0;
    uint64_t zbase = 0;
    std::vector<double> re, im;  // kind 1: 2^nbits | kind 2: (1 + n) * 2^nbits (A then D[r]) | kind 3: 1 + n + n * n
    size_t block_bytes(bool real) const {  // bytes of the occupation block of kinds 2 / 3 / 4 (16-byte header + tables)
        return kind >= 2 ? 16 + re.size() * 8 * (real ? 1 : 2) : 0;
    }
};

static inline uint32_t enum_slot(uint64_t v, uint32_t mult) {
    return ((((uint32_t)(v >> 32)) * ENUM_FOLD + (uint32_t)v) * mult) >> 29;
}

static void analyse_group(int n, uint64_t xy, const int64_t *yz, const double *wre, const double *wim, int num, GroupInfo &out) {
    const uint64_t EVEN = 0x5555555555555555ULL;
    out = GroupInfo();
    if (num < 1) return;
    if (xy == 0) {  // diagonal
        for (int t = 0; t < num; ++t)
            if (__builtin_popcountll((uint64_t)yz[t]) > 2) return;
        out.kind = 3;
        out.re.assign((size_t)1 + n + (size_t)n * n, 0.0);
        out.im.assign(out.re.size(), 0.0);
        for (int pass = 0; pass < 2; ++pass) {
            const double *w = pass ? wim : wre;
            std::vector<double> &tab = pass ? out.im : out.re;
            double *K = &tab[0], *a = &tab[1], *b = &tab[1 + n];
            for (int t = 0; t < num; ++t) {
                const uint64_t z = (uint64_t)yz[t];
                const int pc = __builtin_popcountll(z);
                *K += w[t];
                if (pc == 1) a[__builtin_ctzll(z)] += -2.0 * w[t];
                if (pc == 2) {
                    const int j = __builtin_ctzll(z), i = 63 - __builtin_clzll(z);  // j < i
                    a[i] += -2.0 * w[t];
                    a[j] += -2.0 * w[t];
                    b[(size_t)i * n + j] += 4.0 * w[t];
                }
            }
        }
        return;
    }
    const int ka = __builtin_popcountll(xy & EVEN), kb = __builtin_popcountll(xy & ~EVEN);
    if (ka + kb > 4 || (ka & 1) || (kb & 1)) return;
    // occupation patterns of x' on the mask's positions that a sample of the sector can produce: exactly half of the alpha
    // positions and half of the beta positions occupied
    std::vector<int> pos;
    for (int b = 0; b < 64; ++b)
        if ((xy >> b) & 1) pos.push_back(b);
    std::vector<uint64_t> pats;
    for (int m = 0; m < (1 << pos.size()); ++m) {
        uint64_t b = 0;
        for (size_t i = 0; i < pos.size(); ++i)
            if ((m >> i) & 1) b |= 1ULL << pos[i];
        if (__builtin_popcountll(b & EVEN) * 2 == ka && __builtin_popcountll(b & ~EVEN) * 2 == kb) pats.push_back(b);
    }
    // common Z part outside the mask
    uint64_t zbase = (uint64_t)yz[0] & ~xy;
    int kind = 1;
    for (int t = 0; t < num && kind == 1; ++t)
        if (((uint64_t)yz[t] & ~xy) != zbase) kind = 0;
    if (kind == 0) {  // one extra position at most, relative to E_0 or to E_0 with one position toggled
        const uint64_t e0 = (uint64_t)yz[0] & ~xy;
        for (int c = -1; c < n && kind == 0; ++c) {
            if (c >= 0 && ((xy >> c) & 1)) continue;
            const uint64_t cand = c < 0 ? e0 : e0 ^ (1ULL << c);
            bool good = true;
            for (int t = 0; t < num && good; ++t) good = __builtin_popcountll(((uint64_t)yz[t] & ~xy) ^ cand) <= 1;
            if (good) { kind = 2; zbase = cand; }
        }
        if (kind == 0) return;
    }
    int nbits = 0;
    while ((size_t)(1 << nbits) < pats.size()) ++nbits;
    if (nbits == 0) nbits = 1;
    // multiplier search: a deterministic sequence of odd candidates; the patterns must land on distinct slots < 2^nbits
    uint32_t mult = 0;
    for (; nbits <= 3; ++nbits) {
        uint32_t cand = 0x2545F491u;
        for (int tries = 0; tries < 200000 && !mult; ++tries) {
            cand = cand * 0x9E3779B1u + 0x7F4A7C15u;
            const uint32_t m = cand | 1u;
            uint32_t seen = 0;
            bool good = true;
            for (uint64_t b : pats) {
                const uint32_t sl = enum_slot(b, m);
                if (sl >= (1u << nbits) || ((seen >> sl) & 1u)) { good = false; break; }
                seen |= 1u << sl;
            }
            if (good) mult = m;
        }
        if (mult) break;
    }
    if (!mult) return;  // kind 0
    uint64_t W = 0;  // XOR over the mask's positions i of the bits below i
    for (int i = 0; i < 64; ++i)
        if ((xy >> i) & 1) W ^= (i == 0 ? 0ull : (~0ull >> (64 - i)));
    const bool derived = ((zbase ^ W) & ~xy) == 0;
    if (kind == 1 && !derived) kind = 4;
    out.kind = kind;
    out.nbits = nbits;
    out.mult = mult;
    out.zbase = zbase;
    const size_t ns = (size_t)1 << nbits;
    const size_t rows = kind == 2 ? (size_t)1 + n : 1;  // row 0: G or A; row 1 + r: D[r]
    out.re.assign(rows * ns, 0.0);
    out.im.assign(rows * ns, 0.0);
    for (uint64_t b : pats) {
        const uint32_t sl = enum_slot(b, mult);
        for (int t = 0; t < num; ++t) {
            double sgn = (__builtin_popcountll(b & (uint64_t)yz[t] & xy) & 1) ? -1.0 : 1.0;
            // kind 1: the kernel's sign is parity(x & W) instead of parity(x & zbase); the difference lives on the mask's
            // positions, where x = b ^ xy
            if (kind == 1 && (__builtin_popcountll((b ^ xy) & xy & W) & 1)) sgn = -sgn;
            const uint64_t extra = ((uint64_t)yz[t] & ~xy) ^ zbase;  // 0 or one bit
            out.re[sl] += sgn * wre[t];
            out.im[sl] += sgn * wim[t];
            if (extra) {  // w (-1)^{n_r} = w - 2 w n_r
                const int r = __builtin_ctzll(extra);
                out.re[(size_t)(1 + r) * ns + sl] += -2.0 * sgn * wre[t];
                out.im[(size_t)(1 + r) * ns + sl] += -2.0 * sgn * wim[t];
            }
        }
    }
}

static void build_enum_tiles(int n, const std::vector<uint64_t> &xy, const std::vector<int2> &grp, int64_t U_pad,
                             const int64_t *h_yz, const std::vector<double> &wre, const std::vector<double> &wim, bool real,
                             HostEnum &out) {
    std::vector<GroupInfo> info((size_t)U_pad);
    std::vector<size_t> mask_bytes((size_t)U_pad);
    for (int64_t u = 0; u < U_pad; ++u) {
        analyse_group(n, xy[u], h_yz + grp[u].x, wre.data() + grp[u].x, wim.data() + grp[u].x, grp[u].y, info[u]);
        mask_bytes[u] = 17 + (info[u].kind == 1 ? info[u].re.size() * 8 * (real ? 1 : 2) : info[u].block_bytes(real));
    }
    const int64_t nblocks = U_pad / 32;  // one block = one bitmap word = 32 masks
    std::vector<size_t> block_bytes((size_t)nblocks);
    for (int64_t b = 0; b < nblocks; ++b) {
        size_t bytes = 0;
        for (int64_t u = b * 32; u < b * 32 + 32; ++u) bytes += mask_bytes[u];
        block_bytes[b] = bytes;
    }
    // Tiles of a whole number of expansion steps (ENUM_STEP_WORDS bitmap words) whenever that many words fit: a step that is
    // only partly filled costs the emit kernel as much as a full one.
    const size_t budget = ENUM_TILE_MAX - 256;
    int64_t b = 0;
    while (b < nblocks) {
        const int64_t b0 = b;
        size_t bytes = 0;
        int64_t fit = b0;  // one past the last block that fits
        while (fit < nblocks && bytes + block_bytes[fit] <= budget) bytes += block_bytes[fit++];
        if (fit == b0) { out.ok = false; return; }  // a single word of masks does not fit (a diagonal block of a huge n)
        int64_t take = fit - b0;
        if (fit < nblocks && take > ENUM_STEP_WORDS) take -= take % ENUM_STEP_WORDS;
        b = b0 + take;
        EnumTile tile{};
        tile.u0 = (uint32_t)(b0 * 32);
        tile.n_masks = (uint32_t)(take * 32);
        tile.word0 = (uint32_t)b0;
        tile.n_words = (uint32_t)take;
        const size_t desc_off = (size_t)tile.n_masks * 16;  // (no separate descriptor section any more: kept for the directory)
        const size_t tab_off = desc_off;
        struct Rec { uint64_t xy; uint32_t mult, off; };
        std::vector<Rec> rec(tile.n_masks);
        std::vector<double> tab_re, tab_im;
        for (uint32_t k = 0; k < tile.n_masks; ++k) {  // pattern tables first: [re of all][im of all]
            const GroupInfo &gi = info[(size_t)tile.u0 + k];
            rec[k] = Rec{xy[(size_t)tile.u0 + k], 0u, 0u};
            if (gi.kind == 1) {
                rec[k].mult = gi.mult;
                rec[k].off = (uint32_t)(tab_off + tab_re.size() * 8);
                tab_re.insert(tab_re.end(), gi.re.begin(), gi.re.end());
                if (!real) tab_im.insert(tab_im.end(), gi.im.begin(), gi.im.end());
            }
        }
        tile.n_tab = (uint32_t)tab_re.size();
        tile.desc_off = (uint32_t)desc_off;
        tile.tab_off = (uint32_t)tab_off;
        // occupation blocks of the generic groups: [mult u32][kind | nbits << 8 u32][zbase u64][re tables][im tables when complex]
        std::vector<uint8_t> blocks;
        const size_t blocks_off = tab_off + (tab_re.size() + tab_im.size()) * 8;
        for (uint32_t k = 0; k < tile.n_masks; ++k) {
            const GroupInfo &gi = info[(size_t)tile.u0 + k];
            if (gi.kind == 1 || (size_t)tile.u0 + k >= (size_t)U_pad || grp[(size_t)tile.u0 + k].y == 0) continue;
            tile.n_generic++;
            if (gi.kind < 2) continue;  // kind 0: desc.y = 0 -> term records from global memory
            rec[k].off = (uint32_t)(blocks_off + blocks.size());
            const uint32_t hdr[2] = {gi.mult, (uint32_t)gi.kind | ((uint32_t)gi.nbits << 8)};
            const size_t at = blocks.size();
            blocks.resize(at + gi.block_bytes(real));
            std::memcpy(&blocks[at], hdr, 8);
            std::memcpy(&blocks[at + 8], &gi.zbase, 8);
            std::memcpy(&blocks[at + 16], gi.re.data(), gi.re.size() * 8);
            if (!real) std::memcpy(&blocks[at + 16 + gi.re.size() * 8], gi.im.data(), gi.im.size() * 8);
        }
        auto align16 = [](size_t v) { return (v + 15) / 16 * 16; };
        // bitmap of the generic masks, one word per bitmap word of the tile
        std::vector<uint32_t> genmask(tile.n_words, 0u);
        for (uint32_t k = 0; k < tile.n_masks; ++k)
            if (rec[k].mult == 0u && grp[(size_t)tile.u0 + k].y > 0) genmask[k >> 5] |= 1u << (k & 31u);
        const size_t gen_off = align16(blocks_off + blocks.size());
        tile.gen_off = (uint32_t)gen_off;
        // + 64: slack behind the last table (a slot is always < 8, whatever the table size)
        const size_t nbytes = align16(gen_off + genmask.size() * 4) + 64;
        if (nbytes > ENUM_TILE_MAX) { out.ok = false; return; }
        const size_t off = (out.blob.size() + 127) / 128 * 128;
        out.blob.resize(off + nbytes, 0);
        uint8_t *base = out.blob.data() + off;
        static_assert(sizeof(Rec) == 16, "mask record");
        std::memcpy(base, rec.data(), rec.size() * 16);
        if (!tab_re.empty()) std::memcpy(base + tab_off, tab_re.data(), tab_re.size() * 8);
        if (!tab_im.empty()) std::memcpy(base + tab_off + tab_re.size() * 8, tab_im.data(), tab_im.size() * 8);
        if (!blocks.empty()) std::memcpy(base + blocks_off, blocks.data(), blocks.size());
        std::memcpy(base + gen_off, genmask.data(), genmask.size() * 4);
        tile.blob_off = (uint32_t)off;
        tile.blob_bytes = (uint32_t)nbytes;
        out.tile_bytes_max = std::max(out.tile_bytes_max, tile.blob_bytes);
        out.tiles.push_back(tile);
    }
}


// ---- A13 estimator: packed partial sums of the MonteCarloEstimator (CLE:48-62) in one pass --------------------------------
// out[5] = [sum w, Re sum w E, Im sum w E, Re sum w E^2, Im sum w E^2], w = |psi|^2 (complex square of E, as the reference's
// variance takes it).  One read of both vectors (32 B per row); per-thread partial sums over a grid-stride walk, lanes and
// warps added by shuffles, one partial per block into the workspace, and the block that finishes last adds the partials in
// block order - a fixed summation order for a given n, so two calls return the same bits.
constexpr int STATS_THREADS = 256;
__global__ void __launch_bounds__(STATS_THREADS)
energy_stats_kernel(const double2 *__restrict__ eloc, const double2 *__restrict__ amps, int64_t n, double *__restrict__ partials,
                    unsigned int *__restrict__ done, double *__restrict__ out) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * STATS_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * STATS_THREADS) {
        const double2 e = eloc[i], a = amps[i];
        const double w = a.x * a.x + a.y * a.y;
        const double wer = w * e.x, wei = w * e.y;
        acc[0] += w;
        acc[1] += wer;
        acc[2] += wei;
        acc[3] += wer * e.x - wei * e.y;
        acc[4] += wer * e.y + wei * e.x;
    }
    __shared__ double warp_part[STATS_THREADS / 32][5];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], d);
        if (lane == 0) warp_part[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int wv = 0; wv < STATS_THREADS / 32; ++wv) t += warp_part[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * 5 + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 5) {
        __threadfence();
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += reinterpret_cast<volatile double *>(partials)[(size_t)b * 5 + threadIdx.x];
        out[threadIdx.x] = t;
    }
}

}  // namespace anqs

using namespace anqs;

extern "C" {

int anqs_abi_version(void) { return ANQS_ABI_VERSION; }

const char *anqs_last_error(void) { return g_last_error.c_str(); }

int anqs_device_check(int device, int *sm_count, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    ANQS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    ANQS_REQUIRE(prop.major == 10, "libanqs_b200 is built for sm_100a only; device is compute capability " +
                                       std::to_string(prop.major) + "." + std::to_string(prop.minor));
    return 0;
}

int anqs_popcount_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative element count");
    if (n == 0) return 0;
    ANQS_REQUIRE(d_in && d_out, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int sms = sm_count_of_current_device();
    bool aligned = (((uintptr_t)d_in | (uintptr_t)d_out) & 15) == 0;
    int64_t work = aligned ? (n + 1) / 2 : n;
    int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)sms * 16);
    if (aligned)
        popcount_kernel<<<blocks, 256, 0, s>>>(d_in, d_out, n);
    else
        popcount_kernel_unaligned<<<blocks, 256, 0, s>>>(d_in, d_out, n);
    ANQS_LAUNCH_CHECK();
    return 0;
}

size_t anqs_scan_workspace(int64_t n) {
    int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(ntiles + 1) * sizeof(int64_t);
}

static int stats_grid(int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 4 * STATS_THREADS - 1) / (4 * STATS_THREADS), (int64_t)sm_count_of_current_device() * 4));
}

size_t anqs_energy_stats_workspace(int64_t n) { return (size_t)stats_grid(n) * 5 * sizeof(double) + 16; }

int anqs_energy_stats(const double *d_eloc, const double *d_amps, int64_t n, double *d_out5, void *d_work, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative row count");
    ANQS_REQUIRE(d_out5 && d_work, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        ANQS_CUDA(cudaMemsetAsync(d_out5, 0, 5 * sizeof(double), s));
        return 0;
    }
    ANQS_REQUIRE(d_eloc && d_amps, "null pointer");
    const int grid = stats_grid(n);
    unsigned int *done = reinterpret_cast<unsigned int *>(d_work);
    double *partials = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(d_work) + 16);
    ANQS_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned int), s));
    energy_stats_kernel<<<grid, STATS_THREADS, 0, s>>>((const double2 *)d_eloc, (const double2 *)d_amps, n, partials, done, d_out5);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *d_work, void *stream) {
    ANQS_REQUIRE(n >= 0, "negative element count");
    ANQS_REQUIRE(d_out, "null output");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        ANQS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int64_t), s));
        return 0;
    }
    ANQS_REQUIRE(d_in && d_work, "null pointer");
    int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t *tile_sums = (int64_t *)d_work;
    scan_tile_sums<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(d_in, n, tile_sums);
    ANQS_LAUNCH_CHECK();
    scan_of_tile_sums<<<1, SCAN_BLOCK, 0, s>>>(tile_sums, ntiles);
    ANQS_LAUNCH_CHECK();
    scan_apply<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(d_in, n, tile_sums, d_out);
    ANQS_LAUNCH_CHECK();
    return 0;
}

int anqs_tables_create(anqs_tables_t **out, int qubit_num, int64_t U, int64_t T, const int64_t *h_unq_xy,
                       const int64_t *h_yz_num, const int64_t *h_yz_start, const int64_t *h_yz,
                       const double *h_weights) {
    ANQS_REQUIRE(out, "null handle pointer");
    ANQS_REQUIRE(qubit_num >= 1 && qubit_num <= 64, "qubit_num must be in [1, 64] (single-word indices)");
    ANQS_REQUIRE(U >= 1 && T >= U, "need 1 <= U <= T");
    ANQS_REQUIRE(T < (int64_t)1 << 31, "too many terms");
    ANQS_REQUIRE(h_unq_xy && h_yz_num && h_yz_start && h_yz && h_weights, "null table pointer");
    Tables *t = new Tables();
    std::memset(t, 0, sizeof(Tables));
    t->qubit_num = qubit_num;
    t->U = U;
    t->T = T;
    t->U_pad = (U + 1023) / 1024 * 1024;
    t->row_words = t->U_pad / 32;
    ANQS_CUDA(cudaGetDevice(&t->device));

    std::vector<uint64_t> xy(t->U_pad, 0);
    std::vector<uint2> mab(t->U_pad, make_uint2(0, 0));
    std::vector<int2> grp(t->U_pad, make_int2(0, 0));
    int max_group = 0;
    for (int64_t u = 0; u < U; ++u) {
        uint64_t m = (uint64_t)h_unq_xy[u];
        if (u > 0 && !(h_unq_xy[u - 1] < h_unq_xy[u])) {
            delete t;
            ANQS_REQUIRE(false, "unq_xy_masks must be strictly ascending (signed order)");
        }
        if (h_yz_start[u] < 0 || h_yz_num[u] < 0 || h_yz_start[u] + h_yz_num[u] > T) {
            delete t;
            ANQS_REQUIRE(false, "YZ group out of range");
        }
        xy[u] = m;
        mab[u] = make_uint2(compress_even_bits(m), compress_even_bits(m >> 1));
        grp[u] = make_int2((int)h_yz_start[u], (int)h_yz_num[u]);
        max_group = std::max(max_group, (int)h_yz_num[u]);
    }
    t->max_group = max_group;
    t->max_xy_weight = 0;
    for (int64_t u = 0; u < U; ++u) t->max_xy_weight = std::max(t->max_xy_weight, __builtin_popcountll((unsigned long long)h_unq_xy[u]));
    bool real = true;
    std::vector<double> wre(T), wim(T);
    std::vector<ulonglong2> rec(T);
    std::vector<uint64_t> yzd(T);
    for (int64_t k = 0; k < T; ++k) {
        wre[k] = h_weights[2 * k];
        wim[k] = h_weights[2 * k + 1];
        if (wim[k] != 0.0) real = false;
        unsigned long long bits;
        std::memcpy(&bits, &wre[k], 8);
        yzd[k] = deinterleave((uint64_t)h_yz[k]);
        rec[k] = make_ulonglong2((unsigned long long)yzd[k], bits);
    }
    t->weights_real = real ? 1 : 0;

    auto up = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = up((void **)&t->xy, xy.data(), xy.size() * sizeof(uint64_t));
    if (e == cudaSuccess) e = up((void **)&t->mab, mab.data(), mab.size() * sizeof(uint2));
    if (e == cudaSuccess) e = up((void **)&t->grp, grp.data(), grp.size() * sizeof(int2));
    if (e == cudaSuccess) e = up((void **)&t->yz_d, yzd.data(), (size_t)T * sizeof(uint64_t));
    if (e == cudaSuccess) e = up((void **)&t->w_re, wre.data(), (size_t)T * sizeof(double));
    if (e == cudaSuccess && !real) e = up((void **)&t->w_im, wim.data(), (size_t)T * sizeof(double));
    if (e == cudaSuccess && real) e = up((void **)&t->term_real, rec.data(), (size_t)T * sizeof(ulonglong2));
    HostProd prod;
    build_product_layout(mab, U, prod);
    t->n_tiles = (int)prod.tiles.size();
    t->tile_bytes_max = (int)prod.tile_bytes_max;
    if (e == cudaSuccess) e = up((void **)&t->prod_tiles, prod.tiles.data(), prod.tiles.size() * sizeof(ProdTile));
    if (e == cudaSuccess) e = up((void **)&t->prod_blob, prod.blob.data(), prod.blob.size());
    if (e == cudaSuccess) e = up((void **)&t->prod_row_u, prod.row_u.data(), prod.row_u.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = up((void **)&t->prod_mem_u, prod.mem_u.data(), prod.mem_u.size() * sizeof(uint32_t));
    if (e == cudaSuccess) e = up((void **)&t->prod_blob_u, prod.blob_u.data(), prod.blob_u.size());
    if (e == cudaSuccess) e = up((void **)&t->prod_blob_bs, prod.blob_bs.data(), prod.blob_bs.size());
    t->prod_bs_ok = prod.bs_ok ? 1 : 0;
    {   // bit-sliced filter positions
        const uint64_t EVEN = 0x5555555555555555ULL;
        std::vector<uint2> bs((size_t)t->U_pad, make_uint2(0x40404040u, 0x40404040u));
        t->bs_ok = 1;
        auto part = [&](uint64_t bits) -> uint32_t {
            const int k = __builtin_popcountll(bits);
            if (k & 1) return 0x40404040u;   // can never hold half of its positions occupied
            if (k > 4) { t->bs_ok = 0; return 0x40404040u; }
            uint32_t b[4] = {64, 64, 65, 65};
            int i = 0;
            for (int pos = 0; pos < 64; ++pos)
                if ((bits >> pos) & 1) b[i++] = (uint32_t)pos;
            if (k == 2) { b[2] = 64; b[3] = 65; }
            return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        };
        for (int64_t u = 0; u < U; ++u) bs[u] = make_uint2(part(xy[u] & EVEN), part(xy[u] & ~EVEN));
        if (e == cudaSuccess) e = up((void **)&t->bs_pos, bs.data(), bs.size() * sizeof(uint2));
    }
    HostEnum en;
    build_enum_tiles(qubit_num, xy, grp, t->U_pad, h_yz, wre, wim, real, en);
    if (en.ok) {
        t->n_enum_tiles = (int)en.tiles.size();
        t->enum_tile_bytes_max = (int)en.tile_bytes_max;
        if (e == cudaSuccess) e = up((void **)&t->enum_tiles, en.tiles.data(), en.tiles.size() * sizeof(EnumTile));
        if (e == cudaSuccess) e = up((void **)&t->enum_blob, en.blob.data(), en.blob.size());
    }
    if (e == cudaSuccess) {
        t->dev_copy = nullptr;
        Tables *dc = nullptr;
        e = cudaMalloc((void **)&dc, sizeof(Tables));
        if (e == cudaSuccess) {
            t->dev_copy = dc;
            e = cudaMemcpy(dc, t, sizeof(Tables), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        anqs_tables_destroy((anqs_tables_t *)t);
        set_error(std::string("anqs_tables_create: device upload failed: ") + cudaGetErrorString(e));
        return 2;
    }
    *out = (anqs_tables_t *)t;
    return 0;
}

int anqs_tables_destroy(anqs_tables_t *h) {
    if (!h) return 0;
    Tables *t = (Tables *)h;
    cudaFree(t->xy);
    cudaFree(t->mab);
    cudaFree(t->grp);
    cudaFree(t->yz_d);
    cudaFree(t->w_re);
    cudaFree(t->w_im);
    cudaFree(t->term_real);
    cudaFree(t->prod_tiles);
    cudaFree(t->prod_blob);
    cudaFree(t->prod_row_u);
    cudaFree(t->prod_mem_u);
    cudaFree(t->prod_blob_u);
    cudaFree(t->prod_blob_bs);
    cudaFree(t->bs_pos);
    cudaFree(t->dev_copy);
    cudaFree(t->enum_tiles);
    cudaFree(t->enum_blob);
    delete t;
    return 0;
}

int anqs_tables_info(const anqs_tables_t *h, int *qubit_num, int64_t *U, int64_t *T, int *weights_real,
                     int64_t *bitmap_row_words) {
    ANQS_REQUIRE(h, "null handle");
    const Tables *t = (const Tables *)h;
    if (qubit_num) *qubit_num = t->qubit_num;
    if (U) *U = t->U;
    if (T) *T = t->T;
    if (weights_real) *weights_real = t->weights_real;
    if (bitmap_row_words) *bitmap_row_words = t->row_words;
    return 0;
}

}  // extern "C"
