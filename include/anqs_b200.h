/*
 * anqs_b200.h -- C ABI of libanqs_b200.so: the B200 (sm_100a) kernels behind the VMC inner loop of
 * Exferro/anqs_quantum_chemistry (local-energy evaluation, amplitude evaluation, batch sampling).
 *
 * The reference has NO native plugin/FFI seam apart from two CuPy popcount kernels; its seam is the
 * Python method surface of PauliObservable / HilbertSpace / AbstractANQS.  Each entry
 * point below therefore names the reference *method* it stands behind (paths relative to
 * /root/reference/nqs/nqs/):
 *   PO   = stochastic/observables/pauli_observable.py     HS  = base/hilbert_space.py
 *   POPC = utils/custom_popcount/cuda_int64popcount.py    ANQS = stochastic/ansatzes/anqs/abstract_anqs.py
 *   LAP  = stochastic/ansatzes/anqs/log_abs_phase_anqs.py  MLP = stochastic/ansatzes/anqs/mlp.py
 *   QG   = base/qubit_grouping.py
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device, h_* is a HOST pointer;
 *   - all buffers are caller-allocated; the library owns only opaque handles (anqs_tables_t, anqs_made_t);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise;
 *   - return value 0 = success, non-zero = failure with a thread-local message in anqs_last_error();
 *   - occupation bitstrings are packed one per int64 (qubit_num <= 64, HS:53 int_per_idx == 1), bit i =
 *     base_vec[:, i] (HS:130), OpenFermion qubit q at bit qubit_num-1-q (PO:162); complex128 values are
 *     (re, im) pairs of doubles.
 */
#ifndef ANQS_B200_H
#define ANQS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANQS_ABI_VERSION 1

typedef struct anqs_tables anqs_tables_t; /* device-resident Hamiltonian term tables (PO:103-115)   */
typedef struct anqs_made anqs_made_t;     /* device-resident MADE network + symmetry tables           */

int anqs_abi_version(void);
const char *anqs_last_error(void);
/* SM count / compute capability of `device`; fails unless the device is compute capability 10.x. */
int anqs_device_check(int device, int *sm_count, int *cc_major, int *cc_minor);

/* ---- A4  HilbertSpace.popcount / popcount_ (HS:158-192) = cuda_int64_popcount[_] (POPC:34-87) -------
 * d_out[i] = popcount(d_in[i]); d_out may equal d_in (the in-place variant). */
int anqs_popcount_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *stream);

/* ---- A1  Hamiltonian tables (PO:103-115, built by PO:131-142,185-211) --------------------------------
 * Takes the reference's own five tensors (host memory): unq_xy_masks[U] ascending, unq_xy_to_yz_num[U],
 * unq_xy_to_yz_start[U], rearranged_yz[T], rearranged_weights[T] complex128 as 2T doubles.
 * Synchronous. */
int anqs_tables_create(anqs_tables_t **out, int qubit_num, int64_t U, int64_t T,
                       const int64_t *h_unq_xy, const int64_t *h_yz_num, const int64_t *h_yz_start,
                       const int64_t *h_yz, const double *h_weights);
int anqs_tables_destroy(anqs_tables_t *t);
/* weights_real = 1 when every weight has zero imaginary part (true for real-integral molecules). */
int anqs_tables_info(const anqs_tables_t *t, int *qubit_num, int64_t *U, int64_t *T, int *weights_real,
                     int64_t *bitmap_row_words);

/* ---- A2+A3  kernel 1a: candidates + alpha/beta filter (PO:527-567) -----------------------------------
 * For every sample x and every unique XY mask u: x' = x ^ xy[u] is kept iff
 * popcount(x' & 0x5555..) == alpha_num and popcount(x' & ~0x5555..) == beta_num.
 * d_counts[i]  = number of kept candidates of sample i (may be NULL);
 * d_bitmap     = n rows of bitmap_row_words uint32 words, bit (u & 31) of word (u >> 5) set iff mask u
 *                is kept (may be NULL).  Rows are what anqs_k1_emit consumes. */
int anqs_k1_filter(const anqs_tables_t *t, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                   int64_t *d_counts, uint32_t *d_bitmap, void *stream);

/* d_out[0] = 0, d_out[i+1] = d_out[i] + d_in[i]  (n+1 outputs).  d_work: >= anqs_scan_workspace(n) bytes. */
size_t anqs_scan_workspace(int64_t n);
int anqs_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *d_work, void *stream);

/* ---- A2+A3+A6  kernel 1b: emit the connected list with matrix elements (PO:527-567 + PO:256-324) ------
 * Output rows are lexicographic in (dest, xy_ptr) exactly like the reference's post-filter list; row r of
 * sample i lives at d_offsets[i] + r.  Any of d_dest / d_xy_ptr / d_H may be NULL to skip that column.
 *   d_dest[M] int32   : sample (chunk) index            (PO:531-534 dest_as_chunk_ptrs)
 *   d_xprime[M] int64 : x' = x ^ xy[u]                  (PO:540-541 src_as_base_indices)
 *   d_xy_ptr[M] int32 : u                               (PO:535-538 coupling_xy_as_unq_ham_xy_ptrs)
 *   d_H               : H_{x,x'} = sum_t w_t (-1)^popcount(x' & yz_t)   (PO:308-318);
 *                       h_components = 1 -> M doubles (requires weights_real), 2 -> M complex128. */
int anqs_k1_emit(const anqs_tables_t *t, const int64_t *d_samples, int64_t n, const uint32_t *d_bitmap,
                 const int64_t *d_offsets, int32_t *d_dest, int64_t *d_xprime, int32_t *d_xy_ptr,
                 double *d_H, int h_components, void *stream);

/* ---- A6  PauliObservable.compute_matrix_elements (PO:256-324) on an arbitrary (x', xy_ptr) list ------ */
int anqs_matrix_elements(const anqs_tables_t *t, const int64_t *d_xprime, const int64_t *d_xy_ptr, int64_t m,
                         double *d_H /* m complex128 */, void *stream);

/* ---- A5  kernel 2: membership join, HilbertSpace.find_a_in_b (HS:263-284) -----------------------------
 * Open-addressing table in one caller-allocated buffer of anqs_hash_bytes(capacity) bytes:
 * (capacity + 1) 32-byte slots {key, index, amp.re, amp.im} followed by capacity bytes of presence bits
 * (8 per slot; a probe tests one bit first, so most misses cost a single 4-byte load).  Keys are stored
 * de-interleaved (even bits | odd bits << 32).  capacity: power of two >= anqs_hash_capacity(n).
 * d_amps may be NULL. */
int64_t anqs_hash_capacity(int64_t n);
size_t anqs_hash_bytes(int64_t capacity);
int anqs_hash_build(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                    void *stream);
/* d_ptr[i] = position of d_queries[i] in the key array, or -1; d_mask[i] = (d_ptr[i] != -1) (either may be NULL). */
int anqs_hash_probe(const void *d_table, int64_t capacity, const int64_t *d_queries, int64_t m,
                    int64_t *d_ptr, uint8_t *d_mask, void *stream);

/* ---- A2+A3+A5+A6+A7  fused sample-aware local energy --------------------------------------------------
 * PauliObservable.compute_var_local_energy_proxy(coupling_method='ham') (PO:396-487, non-symmetric branch):
 *   E_loc[i] = ( sum_{x' in sampled set, x' = x_i ^ xy[u] physical} H_{x_i,x'} psi(x') ) / psi(x_i)
 * for rows [row_start, row_start+row_len) of the batch; the sampled set is the table built by
 * anqs_hash_build(d_samples, d_amps, n_total).  Nothing is materialised. */
int anqs_local_energy_sample_aware(const anqs_tables_t *t, const int64_t *d_samples, const double *d_amps,
                                   int64_t n_total, int64_t row_start, int64_t row_len, const void *d_table,
                                   int64_t capacity, int alpha_num, int beta_num, double *d_eloc, void *stream);

/* ---- A7  scatter of the materialised list (PO:453-478 / PO:1048-1057): E[dest] += H * psi(src) -------
 * d_src_ptr[r] = index of x'_r in the sampled set or -1 (skipped).  Rows must be grouped by dest through
 * d_offsets (CSR).  d_eloc[i] = sum / psi(x_i) when d_amps_dest != NULL, else the raw sum. */
int anqs_accumulate_rows(const int64_t *d_offsets, int64_t n, const int64_t *d_src_ptr, const double *d_H,
                         int h_components, const double *d_src_amps, const double *d_amps_dest, double *d_eloc,
                         int accumulate, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ANQS_B200_H */
