/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's local-energy path.
 *
 * Nothing in the product (anqs_quantum_chemistry_b200/) may link, load or call this file.
 * It is the checker used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors; this
 * restatement is pinned against outputs of the unmodified reference itself, run in the build
 * container through oracle/ref_shim.py, committed as tests/golden/ (npz files) by oracle/make_golden.py.
 *
 * Each function cites the reference lines it restates.  Paths are relative to
 * /root/reference/nqs/nqs/ :
 *   PO = stochastic/observables/pauli_observable.py,  HS = base/hilbert_space.py,
 *   POPC = utils/popcount.py
 * Single-word indices only (int_per_idx == 1, i.e. qubit_num <= 64), which covers every
 * BASELINE.json configuration.
 *
 * Build:  make -C oracle   (gcc -O3 -pthread -shared -fPIC; this image has no libgomp, so the
 *          multi-threaded loops use a small pthreads parallel-for)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ---- minimal pthreads parallel-for (dynamic chunks) ---------------------------------------- */
typedef void (*row_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { row_fn fn; void *ctx; int64_t n, chunk; volatile int64_t *next; } pf_t;
static int g_threads = 0;
void orc_set_num_threads(int t) { g_threads = t; }
int orc_num_threads(void) {
    if (g_threads > 0) return g_threads;
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}
static void *pf_worker(void *arg) {
    pf_t *p = (pf_t *)arg;
    for (;;) {
        int64_t lo = __sync_fetch_and_add(p->next, p->chunk);
        if (lo >= p->n) break;
        int64_t hi = lo + p->chunk < p->n ? lo + p->chunk : p->n;
        p->fn(lo, hi, p->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t chunk, row_fn fn, void *ctx) {
    int nt = orc_num_threads();
    if (nt > 256) nt = 256;
    volatile int64_t next = 0;
    pf_t p = { fn, ctx, n, chunk > 0 ? chunk : 1, &next };
    if (nt <= 1 || n <= chunk) { pf_worker(&p); return; }
    pthread_t th[256];
    for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, pf_worker, &p);
    pf_worker(&p);
    for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
}

/* POPC:24-34 -- the reference's SWAR popcount (popcount_mode='memory_efficient'), restated with the
 * same masks rather than a builtin so the oracle does not depend on compiler intrinsics. */
static inline int64_t swar_popcount64(uint64_t x) {
    x = x - ((x >> 1) & 0x5555555555555555ULL);
    x = (x & 0x3333333333333333ULL) + ((x >> 2) & 0x3333333333333333ULL);
    x = (x + (x >> 4)) & 0x0f0f0f0f0f0f0f0fULL;
    return (int64_t)((x * 0x0101010101010101ULL) >> 56);
}

/* HS:158-192 popcount / popcount_ over a flat int64 array (out may alias in). */
void orc_popcount(const int64_t *in, int64_t *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = swar_popcount64((uint64_t)in[i]);
}

/* ---- PO:131-142 + HS:215-228 + PO:185-211: group Pauli terms by unique XY mask --------------
 * in : xy[T], yz[T] (int64, bit 63 = sign bit exactly as PO:164-167 encodes it), w[2T] (re,im)
 * out: unq_xy[U] ascending in SIGNED order (torch.unique sorts int64 as signed),
 *      inv[T], yz_num[U], yz_start[U] (exclusive cumsum, PO:200-202),
 *      re_yz[T], re_w[2T] with the ORIGINAL term order preserved inside each group (PO:194-197).
 * returns U. */
typedef struct { int64_t key; int64_t idx; } kv_t;
static int cmp_kv(const void *a, const void *b) {
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int64_t orc_build_tables(const int64_t *xy, const int64_t *yz, const double *w, int64_t T,
                         int64_t *unq_xy, int64_t *inv, int64_t *yz_num, int64_t *yz_start,
                         int64_t *re_yz, double *re_w) {
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(T > 0 ? T : 1));
    for (int64_t t = 0; t < T; ++t) { kv[t].key = xy[t]; kv[t].idx = t; }
    qsort(kv, (size_t)T, sizeof(kv_t), cmp_kv);
    int64_t U = 0;
    for (int64_t t = 0; t < T; ++t) {
        if (t == 0 || kv[t].key != kv[t - 1].key) { unq_xy[U] = kv[t].key; yz_num[U] = 0; ++U; }
        inv[kv[t].idx] = U - 1;
        yz_num[U - 1] += 1;
    }
    int64_t acc = 0;
    for (int64_t u = 0; u < U; ++u) { yz_start[u] = acc; acc += yz_num[u]; }
    /* kv is sorted by (key, original index) so walking it fills each group in original order */
    for (int64_t t = 0; t < T; ++t) {
        int64_t src = kv[t].idx;
        re_yz[t] = yz[src];
        re_w[2 * t] = w[2 * src];
        re_w[2 * t + 1] = w[2 * src + 1];
    }
    free(kv);
    return U;
}

/* ---- PO:527-567: candidates x' = x ^ xy[u] for all u, then the alpha/beta filter ------------
 * alpha = even bit positions (mask 0x5555..., PO:553), beta = the complement (PO:554).
 * Output is lexicographic in (dest, xy_ptr), which is what the reference's reshape/tile +
 * boolean-mask compaction produces.  Two calls: out arrays NULL -> returns the count only. */
int64_t orc_candidates_ham(const int64_t *samples, int64_t chunk_start, int64_t chunk_len,
                           const int64_t *unq_xy, int64_t U, int64_t alpha_num, int64_t beta_num,
                           int64_t *dest, int64_t *xprime, int64_t *xy_ptr) {
    const uint64_t A = 0x5555555555555555ULL, B = ~A;
    int64_t m = 0;
    for (int64_t i = 0; i < chunk_len; ++i) {
        uint64_t x = (uint64_t)samples[chunk_start + i];
        for (int64_t u = 0; u < U; ++u) {
            uint64_t xp = x ^ (uint64_t)unq_xy[u];
            if (swar_popcount64(xp & A) != alpha_num) continue;
            if (swar_popcount64(xp & B) != beta_num) continue;
            if (dest) { dest[m] = i; xprime[m] = (int64_t)xp; xy_ptr[m] = u; }
            ++m;
        }
    }
    return m;
}

/* ---- HS:263-284 find_a_in_b: a_in_b mask and pointer into b (-1 when absent) -----------------
 * The reference does cat -> unique(return_inverse) -> scatter_ -> gather; for unique b the result
 * is the position of a[i] in b.  Restated with a sorted copy of b and binary search. */
typedef struct { const int64_t *a; const kv_t *kv; int64_t nb; uint8_t *mask; int64_t *ptr; } fab_ctx;
static void fab_rows(int64_t lo_i, int64_t hi_i, void *vc) {
    fab_ctx *c = (fab_ctx *)vc;
    const kv_t *kv = c->kv; int64_t nb = c->nb;
    for (int64_t i = lo_i; i < hi_i; ++i) {
        int64_t lo = 0, hi = nb;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (kv[mid].key < c->a[i]) lo = mid + 1; else hi = mid; }
        if (lo < nb && kv[lo].key == c->a[i]) {
            /* scatter_ with duplicate keys keeps an arbitrary one; for duplicates take the last,
             * which is what sequential scatter_ on CPU does */
            while (lo + 1 < nb && kv[lo + 1].key == c->a[i]) ++lo;
            c->mask[i] = 1; c->ptr[i] = kv[lo].idx;
        } else { c->mask[i] = 0; c->ptr[i] = -1; }
    }
}
void orc_find_a_in_b(const int64_t *a, int64_t na, const int64_t *b, int64_t nb,
                     uint8_t *mask, int64_t *ptr) {
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(nb > 0 ? nb : 1));
    for (int64_t j = 0; j < nb; ++j) { kv[j].key = b[j]; kv[j].idx = j; }
    qsort(kv, (size_t)nb, sizeof(kv_t), cmp_kv);
    fab_ctx c = { a, kv, nb, mask, ptr };
    parallel_for(na, 4096, fab_rows, &c);
    free(kv);
}

/* ---- PO:256-324 compute_matrix_elements ----------------------------------------------------
 * H[i] = sum_{t in group(xy_ptr[i])} (-1)^{popcount(x'[i] & yz[t])} * w[t]   (PO:308-318),
 * terms visited in table order (scatter_add_ on CPU accumulates sequentially). */
typedef struct { const int64_t *xprime, *xy_ptr, *yz_start, *yz_num, *re_yz; const double *re_w; double *H; } me_ctx;
static void me_rows(int64_t lo, int64_t hi, void *vc) {
    me_ctx *c = (me_ctx *)vc;
    for (int64_t i = lo; i < hi; ++i) {
        uint64_t xp = (uint64_t)c->xprime[i];
        int64_t s = c->yz_start[c->xy_ptr[i]], n = c->yz_num[c->xy_ptr[i]];
        double re = 0.0, im = 0.0;
        for (int64_t t = s; t < s + n; ++t) {
            double sgn = (swar_popcount64(xp & (uint64_t)c->re_yz[t]) & 1) ? -1.0 : 1.0;
            re += sgn * c->re_w[2 * t];
            im += sgn * c->re_w[2 * t + 1];
        }
        c->H[2 * i] = re; c->H[2 * i + 1] = im;
    }
}
void orc_matrix_elements(const int64_t *xprime, const int64_t *xy_ptr, int64_t M,
                         const int64_t *yz_start, const int64_t *yz_num,
                         const int64_t *re_yz, const double *re_w, double *H /* [2M] */) {
    me_ctx c = { xprime, xy_ptr, yz_start, yz_num, re_yz, re_w, H };
    parallel_for(M, 2048, me_rows, &c);
}

/* open-addressing set used only to make the fused baseline below run in reasonable time */
typedef struct { uint64_t *keys; int64_t *vals; uint64_t mask; } map_t;
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}
static map_t map_build(const int64_t *keys, int64_t n) {
    map_t m; uint64_t cap = 16; while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
    m.keys = (uint64_t *)malloc(cap * 8); m.vals = (int64_t *)malloc(cap * 8); m.mask = cap - 1;
    for (uint64_t i = 0; i < cap; ++i) m.vals[i] = -1;
    for (int64_t j = 0; j < n; ++j) {
        uint64_t h = mix64((uint64_t)keys[j]) & m.mask;
        while (m.vals[h] != -1 && m.keys[h] != (uint64_t)keys[j]) h = (h + 1) & m.mask;
        m.keys[h] = (uint64_t)keys[j]; m.vals[h] = j;
    }
    return m;
}
static inline int64_t map_get(const map_t *m, uint64_t k) {
    uint64_t h = mix64(k) & m->mask;
    while (m->vals[h] != -1) { if (m->keys[h] == k) return m->vals[h]; h = (h + 1) & m->mask; }
    return -1;
}

/* ---- PO:396-487 compute_var_local_energy_proxy, coupling_method='ham' (non-symmetric branch,
 * PO:469-473) -- the default sample-aware local energy:
 *   E_loc[i] = ( sum_{x' sampled, coupled} H_{x_i,x'} psi(x') ) / psi(x_i)
 * Fused per destination row (no materialisation); mathematically the reference's
 * candidates -> filter -> find_a_in_b -> matrix elements -> scatter_add_ -> divide chain.
 * amps, eloc: complex128 as (re,im) pairs.  Rows [row_start, row_start+row_len) of the batch are
 * evaluated against the WHOLE batch as the sampled set (chunking does not change the result). */
typedef struct {
    const int64_t *samples; const double *amps; int64_t row_start;
    const int64_t *unq_xy; int64_t U; const int64_t *yz_start, *yz_num, *re_yz; const double *re_w;
    int64_t alpha_num, beta_num; double *eloc; const map_t *map;
} le_ctx;
static void le_rows(int64_t lo, int64_t hi, void *vc) {
    le_ctx *c = (le_ctx *)vc;
    const uint64_t A = 0x5555555555555555ULL, B = ~A;
    for (int64_t i = lo; i < hi; ++i) {
        uint64_t x = (uint64_t)c->samples[c->row_start + i];
        double er = 0.0, ei = 0.0;
        for (int64_t u = 0; u < c->U; ++u) {
            uint64_t xp = x ^ (uint64_t)c->unq_xy[u];
            if (swar_popcount64(xp & A) != c->alpha_num) continue;
            if (swar_popcount64(xp & B) != c->beta_num) continue;
            int64_t j = map_get(c->map, xp);
            if (j < 0) continue;
            double hr = 0.0, hi_ = 0.0;
            for (int64_t t = c->yz_start[u]; t < c->yz_start[u] + c->yz_num[u]; ++t) {
                double sgn = (swar_popcount64(xp & (uint64_t)c->re_yz[t]) & 1) ? -1.0 : 1.0;
                hr += sgn * c->re_w[2 * t]; hi_ += sgn * c->re_w[2 * t + 1];
            }
            double ar = c->amps[2 * j], ai = c->amps[2 * j + 1];
            er += hr * ar - hi_ * ai; ei += hr * ai + hi_ * ar;
        }
        double dr = c->amps[2 * (c->row_start + i)], di = c->amps[2 * (c->row_start + i) + 1];
        double den = dr * dr + di * di;
        c->eloc[2 * i] = (er * dr + ei * di) / den;
        c->eloc[2 * i + 1] = (ei * dr - er * di) / den;
    }
}
/* The sampled-set map as an object, so that a caller evaluating many row windows against one sampled set builds it
 * once (the GPU arm builds its table once per batch too). */
void *orc_map_create(const int64_t *samples, int64_t N) {
    map_t *m = (map_t *)malloc(sizeof(map_t));
    *m = map_build(samples, N);
    return m;
}
void orc_map_free(void *vm) {
    map_t *m = (map_t *)vm;
    if (!m) return;
    free(m->keys); free(m->vals); free(m);
}
void orc_local_energy_sample_aware_map(const int64_t *samples, const double *amps, int64_t N, const void *map,
                                       int64_t row_start, int64_t row_len,
                                       const int64_t *unq_xy, int64_t U,
                                       const int64_t *yz_start, const int64_t *yz_num,
                                       const int64_t *re_yz, const double *re_w,
                                       int64_t alpha_num, int64_t beta_num, double *eloc /* [2*row_len] */) {
    (void)N;
    le_ctx c = { samples, amps, row_start, unq_xy, U, yz_start, yz_num, re_yz, re_w, alpha_num, beta_num, eloc, (const map_t *)map };
    parallel_for(row_len, 8, le_rows, &c);
}
void orc_local_energy_sample_aware(const int64_t *samples, const double *amps, int64_t N,
                                   int64_t row_start, int64_t row_len,
                                   const int64_t *unq_xy, int64_t U,
                                   const int64_t *yz_start, const int64_t *yz_num,
                                   const int64_t *re_yz, const double *re_w,
                                   int64_t alpha_num, int64_t beta_num, double *eloc /* [2*row_len] */) {
    map_t map = map_build(samples, N);
    orc_local_energy_sample_aware_map(samples, amps, N, &map, row_start, row_len, unq_xy, U, yz_start, yz_num, re_yz, re_w,
                                      alpha_num, beta_num, eloc);
    free(map.keys); free(map.vals);
}

/* ---- HS:239-261 sort_base_idx: ascending in UNSIGNED order (negatives moved to the end) ------ */
static int cmp_u64(const void *a, const void *b) {
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    uint64_t kx = (uint64_t)x->key, ky = (uint64_t)y->key;
    if (kx != ky) return kx < ky ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
void orc_sort_base_idx(const int64_t *in, int64_t n, int64_t *sorted, int64_t *perm) {
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t j = 0; j < n; ++j) { kv[j].key = in[j]; kv[j].idx = j; }
    qsort(kv, (size_t)n, sizeof(kv_t), cmp_u64);
    for (int64_t j = 0; j < n; ++j) { sorted[j] = kv[j].key; perm[j] = kv[j].idx; }
    free(kv);
}

/* ---- HS:215-228 compute_unique_indices (single word): sorted (signed) unique + inverse -------- */
int64_t orc_unique(const int64_t *in, int64_t n, int64_t *unq, int64_t *inv) {
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t j = 0; j < n; ++j) { kv[j].key = in[j]; kv[j].idx = j; }
    qsort(kv, (size_t)n, sizeof(kv_t), cmp_kv);
    int64_t U = 0;
    for (int64_t j = 0; j < n; ++j) {
        if (j == 0 || kv[j].key != kv[j - 1].key) unq[U++] = kv[j].key;
        inv[kv[j].idx] = U - 1;
    }
    free(kv);
    return U;
}
