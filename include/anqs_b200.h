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
 *   - all buffers are caller-allocated; the library owns only the opaque anqs_tables_t handle;
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

/* ---- A13  packed partial sums of the energy estimator (MonteCarloEstimator with theoretical frequencies, CLE:48-62,
 * 107-113) in one pass over the local energies and the amplitudes of the same rows (both complex128, n rows):
 * d_out5 = [sum w, Re sum w E, Im sum w E, Re sum w E^2, Im sum w E^2], w = |psi|^2 (E^2 = the complex square, as the
 * reference's variance takes it).  mean = sum w E / sum w, var = sum w E^2 / sum w - mean^2; across GPUs the five sums are what
 * one all-reduce adds.  Fixed summation order for a given n.  d_work: >= anqs_energy_stats_workspace(n) bytes, 16-byte aligned. */
size_t anqs_energy_stats_workspace(int64_t n);
int anqs_energy_stats(const double *d_eloc, const double *d_amps, int64_t n, double *d_out5, void *d_work, void *stream);

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

/* ---- A2+A3+A6  kernel 1, tiled variant: the same ordered list at HBM-write speed (PO:527-567 + PO:256-324) ---------
 * Same bitmap, counts and output rows as anqs_k1_filter / anqs_k1_emit (the two pairs are interchangeable bit for bit); the
 * masks are cut into "enumeration tiles" (contiguous mask ranges with the matrix-element tables of their YZ groups) that a
 * CTA keeps resident in shared memory, so matrix elements never go to L2.  anqs_k1_enum_tiles() == 0 means the table does
 * not fit the tiled layout (bitmap rows too long for shared memory): use the pair above.
 * d_work: anqs_k1_enum_workspace(t, n) bytes, 128-byte aligned, written by the filter (the rank of every tile's first
 * connection inside each sample) and consumed by the emit; both calls must be ordered on the same stream.
 * d_counts and d_bitmap are mandatory here; d_offsets = exclusive scan of d_counts (anqs_exclusive_scan_i64). */
int anqs_k1_enum_tiles(const anqs_tables_t *t);
size_t anqs_k1_enum_workspace(const anqs_tables_t *t, int64_t n);
int anqs_k1_enum_filter(const anqs_tables_t *t, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                        int64_t *d_counts, uint32_t *d_bitmap, void *d_work, void *stream);
/* Same with the implementation picked by an ARGUMENT (no process-global state; safe from any number of host threads): the
 * filter has a bit-sliced implementation (32 samples per lane operation; needs every spin part of every mask to have
 * weight <= 4) and a product-layout one (any table).  variant 0 = the bit-sliced one when it applies (what the call above
 * does), 1 = the product-layout one.  Both produce the same bytes. */
int anqs_k1_enum_filter_variant(const anqs_tables_t *t, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                                int64_t *d_counts, uint32_t *d_bitmap, void *d_work, int variant, void *stream);
int anqs_k1_enum_emit(const anqs_tables_t *t, const int64_t *d_samples, int64_t n, int alpha_num, int beta_num,
                      const uint32_t *d_bitmap, const int64_t *d_offsets, void *d_work, int32_t *d_dest, int64_t *d_xprime, int32_t *d_xy_ptr,
                      double *d_H, int h_components, void *stream);

/* ---- A6  PauliObservable.compute_matrix_elements (PO:256-324) on an arbitrary (x', xy_ptr) list ------ */
int anqs_matrix_elements(const anqs_tables_t *t, const int64_t *d_xprime, const int64_t *d_xy_ptr, int64_t m,
                         double *d_H /* m complex128 */, void *stream);

/* ---- A5  kernel 2: membership join, HilbertSpace.find_a_in_b (HS:263-284) -----------------------------
 * Open-addressing table in one caller-allocated, 128-byte aligned buffer of anqs_hash_bytes(capacity) bytes:
 * capacity 32-byte slots {key, index, amp.re, amp.im}, one dedicated slot for the all-ones key, a 96-byte header, padding
 * to the next 8 KB address boundary, and a presence filter of 2*capacity bytes (two bits of one 32-bit word per key; the
 * 128-byte line is chosen by a GF(2)-linear hash of the
 * alpha half of the key, so all the candidates of one sample that share the alpha part of their mask test the same
 * line).  Keys are stored de-interleaved (even bits | odd bits << 32).  capacity: power of two >=
 * anqs_hash_capacity(n).  d_amps may be NULL.  The build is stream-ordered (no host synchronisation); it picks the
 * number of "spread bits" G (keys sharing an alpha half are spread over 2^G lines) from the line occupancy. */
int64_t anqs_hash_capacity(int64_t n);
size_t anqs_hash_bytes(int64_t capacity);
int anqs_hash_build(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                    void *stream);
/* Same with G forced to spread_bits (0..6) instead of chosen from the data. */
int anqs_hash_build_spread(const int64_t *d_keys, const double *d_amps, int64_t n, void *d_table, int64_t capacity,
                           int spread_bits, void *stream);
/* Reads back G and overloaded_keys[g] (g = 0..6) = keys living in lines that hold more than 128 << g keys.
 * Synchronises the stream. */
int anqs_hash_filter_info(const void *d_table, int64_t capacity, int *spread_bits, int64_t *overloaded_keys, void *stream);
/* d_ptr[i] = position of d_queries[i] in the key array, or -1; d_mask[i] = (d_ptr[i] != -1) (either may be NULL). */
int anqs_hash_probe(const void *d_table, int64_t capacity, const int64_t *d_queries, int64_t m,
                    int64_t *d_ptr, uint8_t *d_mask, void *stream);

/* ---- A5 / A12  ordering: sort, unique, top-k (HS:215-261, ANQS:733, PO:1016-1040) -------------------------------------
 * The reference calls torch.sort / torch.unique; these are hand-written radix kernels (k2_sort.cu).  All stream-ordered, no
 * host synchronisation; d_work: the matching *_workspace() bytes, 256-byte aligned.
 *
 * anqs_sort_pairs_u64: stable ascending sort of (key, payload) pairs by the bits [begin_bit, end_bit) of
 * transform(key) ^ xor_mask, transform = identity (key_kind 0) or the order-preserving map of IEEE doubles (key_kind 1).
 * d_vals_in == NULL means payload = position (the permutation comes out in d_vals_out).  xor_mask = 0: unsigned ascending
 * (HilbertSpace.sort_base_idx, HS:239-261); 1 << 63: signed ascending; ~0: descending (ties keep their order: the same rows
 * as torch.sort(descending=True, stable=True)).  Keys come out untransformed.  Not in place. */
size_t anqs_sort_workspace(int64_t n);
int anqs_sort_pairs_u64(const uint64_t *d_keys_in, const int64_t *d_vals_in, uint64_t *d_keys_out, int64_t *d_vals_out, int64_t n,
                        int begin_bit, int end_bit, int key_kind, uint64_t xor_mask, void *d_work, void *stream);
/* HilbertSpace.compute_unique_indices (HS:215-228) for single-word indices: d_unq[0 .. *d_n_unique) = the distinct values of
 * d_in in SIGNED ascending order (torch.unique), d_inv[i] = position of d_in[i] in d_unq (may be NULL).  end_bit: bits at and
 * above it are equal in all inputs (qubit_num, or 64). d_unq must hold n values; *d_n_unique is a device scalar. */
size_t anqs_unique_workspace(int64_t n);
int anqs_unique_i64(const int64_t *d_in, int64_t n, int end_bit, int64_t *d_unq, int64_t *d_inv, int64_t *d_n_unique, void *d_work,
                    void *stream);
/* The k largest of n doubles.  sorted != 0: descending, ties by position - d_top_vals / d_top_idx = the first k rows of
 * torch.sort(vals, descending=True, stable=True) (ANQS:733 keeps exactly those).  sorted == 0: the same SET in position order
 * (ties at the cut still go to the lower positions), which is all a stochastic-beam level needs when its draws are keyed by
 * the row's own identity; saves the final sort.  An 8-bit radix select finds the k-th value from histograms alone (one launch
 * per digit, the last block of each extends the prefix), the survivors are compacted in order and only they are sorted.
 * -inf (masked children) rank last; NaNs are not expected. */
size_t anqs_topk_workspace(int64_t n, int64_t k);
int anqs_topk_f64(const double *d_vals, int64_t n, int64_t k, int sorted, double *d_top_vals, int64_t *d_top_idx, void *d_work,
                  void *stream);

/* ---- A2+A3+A5+A6+A7  fused sample-aware local energy --------------------------------------------------
 * PauliObservable.compute_var_local_energy_proxy(coupling_method='ham') (PO:396-487, non-symmetric branch):
 *   E_loc[i] = ( sum_{x' in sampled set, x' = x_i ^ xy[u] physical} H_{x_i,x'} psi(x') ) / psi(x_i)
 * for rows [row_start, row_start+row_len) of the batch; the sampled set is the table built by
 * anqs_hash_build(d_samples, d_amps, n_total).  Nothing is materialised. */
int anqs_local_energy_sample_aware(const anqs_tables_t *t, const int64_t *d_samples, const double *d_amps,
                                   int64_t n_total, int64_t row_start, int64_t row_len, const void *d_table,
                                   int64_t capacity, int alpha_num, int beta_num, double *d_eloc, void *stream);

/* Same with the kernel picked by an ARGUMENT (no process-global state): the call above runs a bit-sliced kernel (a warp per
 * group of 32 samples; needs every spin part of every mask to have weight <= 4) for batches of >= 256 rows per SM and a
 * warp-per-sample kernel otherwise.  variant 0 = that choice, 1 = the warp-per-sample kernel, 2 = the bit-sliced one
 * wherever the table allows it.  All give the same local energies (order of the fp64 additions aside). */
int anqs_local_energy_sample_aware_variant(const anqs_tables_t *t, const int64_t *d_samples, const double *d_amps,
                                           int64_t n_total, int64_t row_start, int64_t row_len, const void *d_table,
                                           int64_t capacity, int alpha_num, int beta_num, double *d_eloc, int variant,
                                           void *stream);

/* ---- pair-join coupling ('trie' / 'all_to_all' of the reference, PO:602-696, TRIE:8-125) --------------------------------
 * The same sample-aware local energies from the coupled PAIRS of the sampled set: N^2 XOR + POPC tests, survivors look
 * x_i ^ x_j up among the unique XY masks.  d_mask_table: anqs_hash_build over the U unique XY masks (keys = the masks, no
 * amplitudes), capacity >= anqs_hash_capacity(U).  Independent of U; the better algorithm while the sampled set is smaller
 * than the mask list.  d_work: anqs_pair_join_workspace(row_len, n_total) bytes.  Same results as the call above. */
size_t anqs_pair_join_workspace(int64_t row_len, int64_t n_total);
int anqs_local_energy_pair_join(const anqs_tables_t *t, const int64_t *d_samples, const double *d_amps, int64_t n_total,
                                int64_t row_start, int64_t row_len, const void *d_mask_table, int64_t mask_capacity, int alpha_num,
                                int beta_num, double *d_eloc, void *d_work, void *stream);

/* ---- A7  scatter of the materialised list (PO:453-478 / PO:1048-1057): E[dest] += H * psi(src) -------
 * d_src_ptr[r] = index of x'_r in the sampled set or -1 (skipped).  Rows must be grouped by dest through
 * d_offsets (CSR).  d_eloc[i] = sum / psi(x_i) when d_amps_dest != NULL, else the raw sum. */
int anqs_accumulate_rows(const int64_t *d_offsets, int64_t n, const int64_t *d_src_ptr, const double *d_H,
                         int h_components, const double *d_src_amps, const double *d_amps_dest, double *d_eloc,
                         int accumulate, void *stream);

/* ---- A9 + A10  kernel 3: MADE wave function (ANQS:309-485, LAP:14-163, MLP:102-246, QG:99-213, MSK:62-167) ----
 * Plain-data description of a LogAbsPhaseANQS in MADE mode.  Weight pointers are DEVICE pointers to the
 * nn.Linear tensors of log_abs_subnet.layers[l] / phase_subnet.layers[l] (row-major [out][in], already multiplied
 * by the MADE masks as MLP:230-233 does on every forward); biases may be NULL.  depth = number of hidden layers
 * (MLPConfig.depth, default 2), all of width `width` (64), tanh activations, identity on the output layer
 * (MLP:144-148), residual adds on hidden layers 1..depth-1 when use_res (MLP:237-239).
 * sym[s] = {kind, plus_mask, minus_mask, ordinal_mul, ordinal_add, ordinal_div, base, start_eig} describes
 * symmetry s of the LocallyDecomposableMasker: kind 0 (additive) eig = start + popcount(prefix & plus) -
 * popcount(prefix & minus); kind 1 (multiplicative) eig = start * (-1)^popcount(prefix & plus);
 * memo_idx = sum_s ((eig*mul + add) floordiv div) * base (MSK:67-73).
 * cont_mask[q * memo_size + memo_idx] has bit d set iff outcome d of qudit q keeps the prefix physical
 * (QubitGrouping.qudit_idx2cont_mask_mul_table, QG:99-108). */
typedef struct {
    int32_t qubit_num, qudit_num, max_qudit_dim, depth, width, use_res, subtract_mean, sym_num;
    int32_t qudit_starts[65];  /* qudit_starts[qudit_num] == qubit_num */
    uint8_t du[64];            /* 1: local sampling strategy 'DU' for this qudit = all-ones mask (ANQS:417-418) */
    int64_t sym[8][8];
    const double *w_abs[5], *b_abs[5], *w_phase[5], *b_phase[5];
    const uint64_t *cont_mask;
    int64_t memo_size;
} anqs_made_desc_t;

/* AbstractANQS.log_psi (ANQS:407-481) for packed configurations: d_log_psi[i] = (log|psi|, arg psi) as
 * complex128.  Optional training buffers (NULL to skip): d_save_h[2][depth][n][width] = hidden activations of
 * the (log-abs, phase) networks, d_save_p[n][qudit_num][max_qudit_dim] = masked conditional probabilities. */
int anqs_made_log_psi(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi,
                      double *d_save_h, double *d_save_p, void *stream);
/* LogAbsPhaseANQS.cond_log_abs (LAP:105-163) of qudit `qudit_idx` for n packed prefixes (bits at and above
 * qudit_starts[qudit_idx] are ignored): d_cond[n][max_qudit_dim], -inf where the continuation is masked. */
int anqs_made_cond_log_abs(const anqs_made_desc_t *desc, int qudit_idx, const int64_t *d_prefix, int64_t n,
                           double *d_cond, void *stream);

/* Backward of anqs_made_log_psi: the per-sample chain of d log psi / d parameters for both sub-networks, from the
 * activations (d_save_h) and conditional probabilities (d_save_p) the forward call saved and the upstream gradient
 * d_grad_out[n] (complex128: .re multiplies d log|psi|, .im multiplies d arg psi).  Writes the operands of the batch
 * reductions, which are plain GEMMs / column sums and are left to the caller:
 *   d_dY[2][n][qudit_num * max_qudit_dim]  gradient w.r.t. the output layer's pre-activations (net 0 = log-abs, 1 = phase)
 *   d_da[2][depth][n][width]               gradient w.r.t. the pre-activation of hidden layer l
 *   d_x[n][qubit_num]                      the 1 - 2 bit input encoding
 * grad W_out = dY^T h_last, grad b_out = sum_s dY, grad W_l = da_l^T (h_{l-1} | x), grad b_l = sum_s da_l. */
int anqs_made_backward_chain(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                             const double *d_save_h, const double *d_save_p, double *d_dY, double *d_da, double *d_x,
                             void *stream);
/* The same chain without the phase network's half of d_dY: d_dY_abs[n][qudit_num * max_qudit_dim] is the log-abs
 * network's output-layer gradient only.  The phase network's output-layer gradient has one non-zero per sample and qudit
 * (arg psi = pi * sum_q y[q, chosen_q], LAP:97, ANQS:450-454), so its weight / bias gradient
 *   d_gW[qudit_num * max_qudit_dim][width] (+)= sum_s pi g_s.im h_s   into row  q * max_qudit_dim + chosen_q(s),   d_gb likewise,
 * is a row scatter (anqs_made_phase_output_grad; d_h_last[n][width] = the phase network's last hidden activations, i.e.
 * d_save_h[1][depth - 1]) instead of a dense product with 63/64 zeros.  Deterministic (fixed summation order);
 * d_work: anqs_made_phase_output_workspace(desc) bytes; d_gb may be null. */
int anqs_made_backward_chain_abs(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                 const double *d_save_h, const double *d_save_p, double *d_dY_abs, double *d_da, double *d_x,
                                 void *stream);
int64_t anqs_made_phase_output_workspace(const anqs_made_desc_t *desc);
int anqs_made_phase_output_grad(const anqs_made_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                                const double *d_h_last, int accumulate, double *d_gW, double *d_gb, void *d_work,
                                int64_t work_bytes, void *stream);

/* ---- A9  kernel 3, NADE mode (ANQS:410-428, LAP:24-42): one (log-abs, phase) MLP pair per qudit --------------------------
 * Same scalar fields as anqs_made_desc_t.  d_ptrs is a DEVICE array of 2 * qudit_num * (depth + 1) * 2 device pointers:
 * entry ((net * qudit_num + q) * (depth + 1) + layer) * 2 + {0: weight [out][in], 1: bias or NULL}; net 0 =
 * log_abs_subnet[q], net 1 = phase_subnet[q]; layer 0 has in = max(1, qudit_starts[q]) inputs (LAP:26), the last layer
 * out = 2^(qubits of qudit q) outputs.  Differences from MADE mode: the mean is subtracted over the qudit's own
 * outcomes (LAP:118-119).  d_save_h: [2][qudit_num][depth][n][width], d_save_p: [n][qudit_num][max_qudit_dim]. */
typedef struct {
    int32_t qubit_num, qudit_num, max_qudit_dim, depth, width, use_res, subtract_mean, sym_num;
    int32_t qudit_starts[65];
    uint8_t du[64];
    int64_t sym[8][8];
    const double *const *ptrs;
    const uint64_t *cont_mask;
    int64_t memo_size;
} anqs_nade_desc_t;
int anqs_nade_log_psi(const anqs_nade_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, double *d_save_h,
                      double *d_save_p, void *stream);
int anqs_nade_cond_log_abs(const anqs_nade_desc_t *desc, int qudit_idx, const int64_t *d_prefix, int64_t n, double *d_cond,
                           void *stream);
/* The reductions over the batch that finish a backward pass (reference: torch's addmm backward under MLP:217-246):
 * for every problem  C[M][N] (+)= A^T B  and, when colsum is not null,  colsum[M] (+)= column sums of A,  with A [K][M]
 * (leading dimension lda) and B [K][N] (ldb) sample-major and N <= 64.  All problems share K.  Deterministic: partial tiles
 * go to `workspace` (anqs_batch_reduce_workspace bytes) and are added in a fixed order.  accumulate != 0 adds to C / colsum. */
typedef struct {
    const double *A, *B;
    double *C, *colsum;
    int32_t lda, ldb, ldc, M, N, reserved;
} anqs_brg_problem_t;
int64_t anqs_batch_reduce_workspace(const anqs_brg_problem_t *problems, int n_problems, int64_t K);
int anqs_batch_reduce_gemm(const anqs_brg_problem_t *problems, int n_problems, int64_t K, int accumulate, void *workspace,
                           int64_t workspace_bytes, void *stream);

/* Backward chain of anqs_nade_log_psi, as anqs_made_backward_chain but per (sub-network, qudit) MLP:
 * d_dY[2][n][qudit_num * max_qudit_dim] (entries beyond the qudit's own outcomes are zero), d_da[2][qudit_num][depth][n][width],
 * d_x[n][qubit_num]; d_save_h / d_save_p as anqs_nade_log_psi saved them. */
int anqs_nade_backward_chain(const anqs_nade_desc_t *desc, const int64_t *d_idx, int64_t n, const double *d_grad_out,
                             const double *d_save_h, const double *d_save_p, double *d_dY, double *d_da, double *d_x, void *stream);

/* NADE mode on the tensor cores (k3_nade_tc.cu): as the MADE tensor-core entry points below, for the per-qudit MLP pairs.
 * anqs_nade_tc_pack must be called again whenever a parameter changes; d_packed: anqs_nade_tc_packed_bytes(desc) bytes,
 * 128-byte aligned.  Inference only, stated tolerance (tests/test_gpu_nade.py). */
size_t anqs_nade_tc_packed_bytes(const anqs_nade_desc_t *desc);
int anqs_nade_tc_pack(const anqs_nade_desc_t *desc, void *d_packed, void *stream);
int anqs_nade_log_psi_tc(const anqs_nade_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n, double *d_log_psi,
                         void *stream);
int anqs_nade_cond_log_abs_tc(const anqs_nade_desc_t *desc, const void *d_packed, int qudit_idx, const int64_t *d_prefix, int64_t n,
                              double *d_cond, void *stream);

/* ---- A9  kernel 3, tensor-core mode: the same two functions with every GEMM on tcgen05 (kind::tf32, fp32 accumulate in
 * TMEM) and fp32 epilogue math.  Inference only (no activations are saved); agreement with the fp64 entry points above
 * is ~1e-3 in log|psi| and ~1e-2 rad in the phase (tf32 products carry 10 mantissa bits), see tests/test_gpu_anqs.py.
 * The weights are first packed into the tensor cores' shared-memory operand layout: anqs_made_tc_pack must be called
 * again whenever a weight changes.  d_packed: anqs_made_tc_packed_bytes(desc) bytes, 128-byte aligned.
 * Requires width == 64 and max_qudit_dim <= 64 like the fp64 kernels. */
size_t anqs_made_tc_packed_bytes(const anqs_made_desc_t *desc);
int anqs_made_tc_pack(const anqs_made_desc_t *desc, void *d_packed, void *stream);
int anqs_made_log_psi_tc(const anqs_made_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n,
                         double *d_log_psi, void *stream);
int anqs_made_cond_log_abs_tc(const anqs_made_desc_t *desc, const void *d_packed, int qudit_idx, const int64_t *d_prefix,
                              int64_t n, double *d_cond, void *stream);

/* ---- A9 (config 3)  kernel 5: autoregressive transformer wave function ------------------------------------------------
 * Architecture of the reference's TransformerMADE (stochastic/ansatzes/legacy/anqs_primitives/made/transformer_made.py:9-48):
 * token embedding [3][dim] (token 2 = BOS) + positional embedding [qubit_num+1][dim]; `depth` post-norm
 * nn.TransformerEncoderLayer blocks (in_proj [3 dim][dim] + bias, out_proj, linear1 / linear2 with dim_feedforward = dim and
 * ReLU, norm1 / norm2); decoder [4][dim] -> (re, im) of outcome 0, (re, im) of outcome 1 per position
 * (legacy/made/real_log_psi_transformer_made.py:42-58).  All pointers are DEVICE pointers to the float64 torch parameters
 * (row-major [out][in]).  sym / cont_mask / memo_size as in anqs_made_desc_t with ONE qubit per qudit: cont_mask[t *
 * memo_size + memo_idx] has bit o set iff outcome o of qubit t keeps the prefix physical.  dim must be 64. */
typedef struct {
    int32_t qubit_num, dim, depth, head_num, sym_num, pad0, pad1, pad2;
    int64_t sym[8][8];
    const double *tok_emb, *pos_emb;
    const double *in_proj_w[4], *in_proj_b[4], *out_proj_w[4], *out_proj_b[4];
    const double *lin1_w[4], *lin1_b[4], *lin2_w[4], *lin2_b[4];
    const double *ln1_w[4], *ln1_b[4], *ln2_w[4], *ln2_b[4];
    const double *dec_w, *dec_b;
    const uint64_t *cont_mask;
    int64_t memo_size;
    double ln_eps;
} anqs_transformer_desc_t;

/* d_log_psi[i] = sum over qubits of the masked, normalised conditional log-amplitude at the chosen outcome (complex128:
 * log|psi|, arg psi); unphysical configurations give (-inf, 0). */
int anqs_transformer_log_psi(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, void *stream);
/* d_cond[i][2] = normalised conditional log|psi| of qubit `qubit_idx` for the two outcomes given the first qubit_idx bits of
 * d_prefix[i] (-inf where the continuation is masked). */
int anqs_transformer_cond_log_abs(const anqs_transformer_desc_t *desc, int qubit_idx, const int64_t *d_prefix, int64_t n,
                                  double *d_cond, void *stream);

/* Gradient of anqs_transformer_log_psi with respect to every parameter (k5_transformer_bwd.cu); replaces autograd through
 * TransformerMADE (legacy/anqs_primitives/made/transformer_made.py:9-48) and the masked normalisation (ANQS:392-405).
 * `grads` mirrors the descriptor's parameter pointers with the destinations of the gradients (same shapes; bias pointers
 * may be null).  d_grad_out[n] complex128 = dLoss/d(log|psi|), dLoss/d(arg psi) per sample, i.e. the gradient is
 * sum_i  grad_out[i].re * d log|psi_i| / d theta + grad_out[i].im * d arg psi_i / d theta.
 * Two ways to call it, with a workspace of anqs_transformer_backward_workspace(desc, n) bytes (256-byte aligned):
 *   saved != 0: anqs_transformer_log_psi_saving(desc, d_idx, n, d_log_psi, d_work, ...) ran before on the same samples and
 *               the same workspace (it computes the same d_log_psi as anqs_transformer_log_psi, bit for bit, and leaves
 *               the activations in d_work); the backward pass starts from them;
 *   saved == 0: the forward pass is recomputed inside; with accumulate != 0 (adds to the destinations) this is how a
 *               caller walks a batch too large for one workspace in chunks of n samples.
 * All sums over the batch run in a fixed order: results are reproducible bit for bit. */
typedef struct {
    double *tok_emb, *pos_emb;
    double *in_proj_w[4], *in_proj_b[4], *out_proj_w[4], *out_proj_b[4];
    double *lin1_w[4], *lin1_b[4], *lin2_w[4], *lin2_b[4];
    double *ln1_w[4], *ln1_b[4], *ln2_w[4], *ln2_b[4];
    double *dec_w, *dec_b;
} anqs_transformer_grads_t;
int64_t anqs_transformer_backward_workspace(const anqs_transformer_desc_t *desc, int64_t n);
int anqs_transformer_log_psi_saving(const anqs_transformer_desc_t *desc, const int64_t *d_idx, int64_t n, double *d_log_psi, void *d_work,
                                    int64_t work_bytes, void *stream);
int anqs_transformer_backward(const anqs_transformer_desc_t *desc, const anqs_transformer_grads_t *grads, const int64_t *d_idx, int64_t n,
                              const double *d_grad_out, void *d_work, int64_t work_bytes, int saved, int accumulate, void *stream);

/* Tensor-core mode of the two functions above: every projection on tcgen05 (kind::tf32, fp32 accumulation in TMEM), attention,
 * LayerNorm and the masked normalisation in fp32.  Inference only; agreement with the fp64 entry points is a stated
 * tolerance (tests/test_gpu_transformer.py).  The parameters are first packed into the tensor cores' operand layout:
 * anqs_transformer_tc_pack must be called again whenever one changes.  d_packed: anqs_transformer_tc_packed_bytes(desc)
 * bytes, 128-byte aligned.  head_num must be 4, 8 or 16. */
size_t anqs_transformer_tc_packed_bytes(const anqs_transformer_desc_t *desc);
int anqs_transformer_tc_pack(const anqs_transformer_desc_t *desc, void *d_packed, void *stream);
int anqs_transformer_log_psi_tc(const anqs_transformer_desc_t *desc, const void *d_packed, const int64_t *d_idx, int64_t n,
                                double *d_log_psi, void *stream);
int anqs_transformer_cond_log_abs_tc(const anqs_transformer_desc_t *desc, const void *d_packed, int qubit_idx, const int64_t *d_prefix,
                                     int64_t n, double *d_cond, void *stream);

/* ---- A11  kernel 4: one level of the count-splitting batch sampler (ANQS:593-662) ----------------------
 * Parents i = 0..n-1 carry a packed prefix, a count (double, exact below 2^53), and a memo index.
 * split:  d_child_counts[i][D] (D = 2^qubits_in_qudit) = exact multinomial split of d_counts[i] with
 *         probabilities softmax(2*d_cond[i][:D]) drawn as qubits_in_qudit rounds of binomials on the cumulative
 *         tree, most significant outcome bit first (ANQS:557-591); d_n_children[i] = number of children that
 *         are allowed by d_cont_mask_q[memo_idx] and have count > 0 (ANQS:653-660).
 *         draw_mode 0: every binomial draw is replaced by rint(n*p) (deterministic, used for parity tests);
 *         draw_mode 1: Philox4x32-10 binomial variates keyed by (seed; key_i, level, round, node) with key_i =
 *         d_rng_keys[i] when given (pass the packed prefixes: the draws then do not depend on how the nodes of a level are
 *         split over launches, ranks or GPUs) else parent_offset + i.
 *         d_single (optional, int8[n]): a parent that carries one sample in draw_mode 1 then gets its child as a byte
 *         (outcome, or -1 when the symmetry table forbids it) INSTEAD of a row of d_child_counts; all other parents get -2
 *         and their dense row.  Pass the same array to emit.  NULL: every parent gets its dense row.
 * emit:   writes the surviving children in (parent, outcome) order at d_offsets[i] (exclusive scan of
 *         d_n_children): prefix | outcome << qudit_start, count, next memo index (QG:99-108 tables as int32). */
int anqs_sampler_split_level(const double *d_cond, int max_qudit_dim, int qubits_in_qudit, const double *d_counts,
                             const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n,
                             int level, int draw_mode, uint64_t seed, int64_t parent_offset, const int64_t *d_rng_keys,
                             double *d_child_counts, int64_t *d_n_children, int8_t *d_single, void *stream);
int anqs_sampler_emit_children(const double *d_child_counts, int qubits_in_qudit, int qudit_start,
                               const int64_t *d_prefix, const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q,
                               const int32_t *d_next_memo_q, int64_t memo_size, int64_t n, const int64_t *d_offsets,
                               const int8_t *d_single, int64_t *d_out_prefix, double *d_out_counts, int32_t *d_out_memo_idx,
                               void *stream);
/* Same with a bound on the output rows: children that would land at or beyond out_capacity are not written.  Lets a caller
 * size the next level from a prediction (the previous call's level sizes) instead of reading every level's size back to the
 * host: it checks all the d_offsets totals once, after the last level, and repeats the call exactly if one exceeded its bound. */
int anqs_sampler_emit_children_capped(const double *d_child_counts, int qubits_in_qudit, int qudit_start,
                                      const int64_t *d_prefix, const int32_t *d_memo_idx, const uint64_t *d_cont_mask_q,
                                      const int32_t *d_next_memo_q, int64_t memo_size, int64_t n, const int64_t *d_offsets,
                                      const int8_t *d_single, int64_t out_capacity, int64_t *d_out_prefix, double *d_out_counts,
                                      int32_t *d_out_memo_idx, void *stream);

/* ---- A12  one level of Gumbel top-k (stochastic beam) sampling (ANQS:676-688, 718-731) -----------------
 * d_out_log_prob[i][D] = d_parent_log_prob[i] + 2*d_cond[i][:D]; d_out_gumbel[i][D] = Gumbel(log_prob)
 * conditioned on max = d_parent_gumbel[i]; masked children get -inf.  Uniforms come from d_uniforms[i][D]
 * when given (parity tests), otherwise from Philox4x32-10 keyed like the split kernel. */
int anqs_sampler_gumbel_level(const double *d_cond, int max_qudit_dim, int qubits_in_qudit,
                              const double *d_parent_log_prob, const double *d_parent_gumbel, const int32_t *d_memo_idx,
                              const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n, int level, uint64_t seed,
                              int64_t parent_offset, const double *d_uniforms, double *d_out_log_prob,
                              double *d_out_gumbel, void *stream);
/* Same with the Philox draws keyed by d_rng_keys[i] (pass the packed prefixes) instead of parent_offset + i: a level then
 * gives the same children whatever the order or position of its rows, so the per-level top-k need not be sorted. */
int anqs_sampler_gumbel_level_keyed(const double *d_cond, int max_qudit_dim, int qubits_in_qudit,
                                    const double *d_parent_log_prob, const double *d_parent_gumbel, const int32_t *d_memo_idx,
                                    const uint64_t *d_cont_mask_q, int64_t memo_size, int64_t n, int level, uint64_t seed,
                                    int64_t parent_offset, const int64_t *d_rng_keys, const double *d_uniforms, double *d_out_log_prob,
                                    double *d_out_gumbel, void *stream);

/* Survivors of one Gumbel top-k level (ANQS:733-776).  d_sorted_idx / d_sorted_gumbel = the level's [n*D] perturbed
 * log-probabilities sorted in descending order (flat index parent * D + outcome; the global sort itself is a library call).
 * Rows r < keep become the nodes of the next level: d_out_prefix[r] = d_prefix[parent] | outcome << qudit_start,
 * d_out_memo_idx[r] = d_next_memo_q[memo_idx[parent] * D + outcome], d_out_log_prob[r] = d_level_log_prob[flat],
 * d_out_gumbel[r] = d_sorted_gumbel[r].  Masked children carry -inf and sort last; *d_n_alive (device int32) = number of
 * rows in front of them.  The caller may keep [0, *d_n_alive) or carry all `keep` rows on: dead rows come out with memo
 * index -1 and log-probability -inf, so every level below masks them again (one host read at the end instead of one per level). */
int anqs_sampler_gumbel_select(const int64_t *d_sorted_idx, const double *d_sorted_gumbel, int64_t keep, int qubits_in_qudit,
                               int qudit_start, const int64_t *d_prefix, const int32_t *d_memo_idx, const int32_t *d_next_memo_q,
                               const double *d_level_log_prob, int64_t *d_out_prefix, int32_t *d_out_memo_idx,
                               double *d_out_log_prob, double *d_out_gumbel, int32_t *d_n_alive, void *stream);

/* The same for a level that was drawn from UNMASKED conditionals (LocalSamplingConfig(masking_depth > 0): strategy 'DU',
 * ANQS:708-709): there the unphysical children compete in the top-k with finite Gumbels and are dropped afterwards
 * (ANQS:804-809).  d_drop_mask_q[memo_size] = the qudit's TRUE continuation-mask words; a kept row whose outcome bit is clear
 * comes out dead (memo index -1, log-probability and Gumbel -inf) and is not counted in *d_n_alive - dead rows may then sit
 * anywhere among the `keep` rows.  d_drop_mask_q = NULL: anqs_sampler_gumbel_select. */
int anqs_sampler_gumbel_select_masked(const int64_t *d_sorted_idx, const double *d_sorted_gumbel, int64_t keep, int qubits_in_qudit,
                                      int qudit_start, const int64_t *d_prefix, const int32_t *d_memo_idx,
                                      const int32_t *d_next_memo_q, const double *d_level_log_prob, const uint64_t *d_drop_mask_q,
                                      int64_t memo_size, int64_t *d_out_prefix, int32_t *d_out_memo_idx, double *d_out_log_prob,
                                      double *d_out_gumbel, int32_t *d_n_alive, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ANQS_B200_H */
