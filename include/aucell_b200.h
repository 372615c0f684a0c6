/*
 * aucell_b200.h -- C-ABI of the B200-native AUCell hot path (libaucell_b200.so).
 *
 * Drop-in boundary for the one data-parallel path of aertslab/pySCENIC this repository replaces:
 *
 *     pyscenic.aucell.create_rankings      reference src/pyscenic/aucell.py:25-51
 *     pyscenic.aucell.derive_auc_threshold reference src/pyscenic/aucell.py:54-72   (per-cell non-zero counts)
 *     pyscenic.aucell.aucell4r/_enrichment reference src/pyscenic/aucell.py:78-182
 *     pyscenic.aucell.aucell               reference src/pyscenic/aucell.py:185-213
 *     ctxcore.recovery.enrichment4cells -> aucs -> auc2d -> weighted_auc1d
 *                                          (third-party, call sites src/pyscenic/aucell.py:95,122-126)
 *
 * The reference is pure Python, so "the reference's FFI for this path" is a ctypes binding; the stub a
 * maintainer would add to pyscenic/aucell.py is shown in INTEGRATION.md.  Everything here is extern "C" with
 * plain pointers and sizes.  No torch types.  There is NO CPU fallback: every entry point needs an sm_100
 * device and fails with AUCELL_ERR_CUDA / AUCELL_ERR_NO_DEVICE otherwise.
 *
 * Division of labour.  The HOST (Python, pyscenic_b200/) owns: the seeded column
 * permutation (numpy legacy RandomState, as pandas' DataFrame.sample draws it), gene-name -> column
 * resolution, the 80 % rule, noweights, rank_cutoff (Python banker's rounding), max_auc, normalize, and
 * DataFrame assembly.  THIS LIBRARY owns everything per cell: ranking with the seeded tie order and the
 * recovery-curve AUC of every regulon.
 *
 * Conventions
 *   - "shuffled position" j of a gene column g: perm[j] == g.  Ties rank in ascending shuffled position, NaN
 *     last, -0.0 == +0.0 (pandas rank(method="first", ascending=False, na_option="bottom") after the column
 *     shuffle, aucell.py:46-51).
 *   - Ranking matrices are uint32, 0-based, laid out [n_cells x n_genes] in SHUFFLED column order (column j
 *     is gene perm[j]) -- exactly the frame create_rankings returns (aucell.py:47 keeps the shuffled order).
 *   - AUC matrices are float64 [n_cells x n_regulons], row-major, regulons in the order given to the plan.
 *   - CSR inputs must be canonical (column indices unique per row); indptr is int64, indices int32.
 *   - All device entry points are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the legacy
 *     default stream).  Pointers named d_* are device pointers owned by the caller, h_* host pointers.
 *   - Return value: 0 on success, a negative AUCELL_ERR_* otherwise; aucell_b200_last_error() gives the
 *     message of the last failure on the calling thread.
 */
#ifndef AUCELL_B200_H
#define AUCELL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AUCELL_B200_ABI_VERSION 4   /* 4: + aucell_csr_is_canonical, aucell_b200_release_staging, aucell_dense_strided_host, aucell_cast_f64_to_f32, aucell_cast_i64_to_i32, aucell_host_dense_to_csr; 3: + aucell_plan_set_fast_path, aucell_plan_fast_stats, aucell_csr16_lut8_f32, aucell_csr_lut8_f32, aucell_count_nonzero_csr_f32, aucell_binarize_*; 2: + aucell_csr16_f32, aucell_rss_f64, aucell_plan_last_host_bytes */

enum {
    AUCELL_OK = 0,
    AUCELL_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, cutoff out of range, ...) */
    AUCELL_ERR_CUDA = -2,      /* a CUDA runtime call or kernel failed                                 */
    AUCELL_ERR_NO_DEVICE = -3, /* no CUDA device / not an sm_100 device                                */
    AUCELL_ERR_UNSUPPORTED = -4, /* shape outside what the kernels support (see aucell_b200_limits)   */
    AUCELL_ERR_NOMEM = -5,
    AUCELL_NOT_SPARSE = 1      /* aucell_dense_strided_host only: nothing computed, use another entry point     */
};

/* opaque: per-call constants of one AUCell job, resident on one device */
typedef struct aucell_plan aucell_plan_t;

/* ---- library ---------------------------------------------------------------------------------------- */
int aucell_b200_abi_version(void);
const char* aucell_b200_last_error(void);
/* device_count < 0 on failure.  cc = 10*major + minor of `device`; sm_count its SM count. */
int aucell_b200_device_info(int device, int* device_count, int* cc, int* sm_count, int64_t* smem_optin_bytes);
/* largest n_genes the shared-memory kernels accept on this build (~600 k); plans beyond it fail with
 * AUCELL_ERR_UNSUPPORTED at launch */
int64_t aucell_b200_max_genes(void);

/* ---- plan: permutation + regulon tables + cutoff (host -> device, once per aucell() call) ------------
 *
 * h_perm[n_genes]          perm[j] = gene column at shuffled position j.  NULL = identity (aucell4r path,
 *                          where the caller already holds a ranking matrix).
 * n_regulons, h_reg_indptr[n_regulons+1], h_reg_cols[nnz], h_reg_weights[nnz]
 *                          regulon CSR over gene COLUMN indices of the expression matrix (the columns
 *                          rnk_mtx.columns.isin(regulon.genes) selects, ctxcore enrichment4cells), fp64 weights
 *                          (regulon[gene]).  h_reg_weights == NULL means every weight is 1.0 (noweights=True,
 *                          aucell.py:124,147-148): the kernels then accumulate in exact integers.
 * h_reg_max_auc[n_regulons] float((rank_cutoff + 1) * weights.sum())  (ctxcore aucs); must be > 0 where valid.
 * h_reg_valid[n_regulons]  0 = fewer than 80 % of the regulon's genes are in the matrix -> column of zeros.
 * rank_cutoff              ctxcore derive_rank_cutoff: int(round(auc_threshold * n_genes)) - 1; only ranks
 *                          strictly below it contribute.  0 <= rank_cutoff <= 2^30 (AUCell itself has
 *                          rank_cutoff < n_genes; ctxcore.recovery.aucs on a column-selected ranking matrix derives
 *                          the cutoff from the whole genome and may exceed the number of columns).
 *                          n_regulons == 0 is allowed (ranking-only plan; rank_cutoff ignored).
 */
int aucell_plan_create(int device, int64_t n_genes, const int32_t* h_perm, int32_t n_regulons,
                       const int64_t* h_reg_indptr, const int32_t* h_reg_cols, const double* h_reg_weights,
                       const double* h_reg_max_auc, const uint8_t* h_reg_valid, int64_t rank_cutoff,
                       aucell_plan_t** out_plan);
int aucell_plan_destroy(aucell_plan_t* plan);

/* ---- create_rankings (aucell.py:25-51): full uint32 ranking rows, shuffled column order --------------
 * d_out_rank [n_cells x n_genes] uint32.  d_out_nnz (nullable) [n_cells] int32 = np.count_nonzero per cell
 * (derive_auc_threshold, aucell.py:67: NaN counts as non-zero, -0.0 does not). */
int aucell_rank_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream);
int aucell_rank_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                          uint32_t* d_out_rank, int32_t* d_out_nnz, void* stream);
int aucell_rank_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const float* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream);
int aucell_rank_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const double* d_data, int64_t n_cells, uint32_t* d_out_rank, int32_t* d_out_nnz,
                        void* stream);

/* ---- aucell (aucell.py:185-213): ranking fused with the AUC of every regulon; the cells x genes ranking
 * matrix is never written.  d_out_auc [n_cells x n_regulons] float64, d_out_nnz nullable as above. -------- */
int aucell_dense_f32(const aucell_plan_t* plan, const float* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream);
int aucell_dense_f64(const aucell_plan_t* plan, const double* d_expr, int64_t n_cells, int64_t row_stride,
                     double* d_out_auc, int32_t* d_out_nnz, void* stream);
int aucell_csr_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const float* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream);
int aucell_csr_f64(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                   const double* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream);
/* CSR with uint16 column indices (plans with n_genes <= 65535 only): 6 instead of 8 bytes per stored entry.  Same
 * contract as aucell_csr_f32; it is what aucell_csr_f32_host uploads after narrowing the caller's int32 indices. */
int aucell_csr16_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const uint16_t* d_indices,
                     const float* d_data, int64_t n_cells, double* d_out_auc, int32_t* d_out_nnz, void* stream);

/* CSR with uint16 column indices and ONE-BYTE value codes: d_codes[e] indexes d_lut, a table of up to 256 float32 values
 * (the distinct values of the matrix; counts and their normalisations have few).  3 instead of 8 bytes per stored entry:
 * what aucell_csr_f32_host uploads while the matrix allows it.  d_vals_scratch (float, indexed like d_codes) receives
 * the decoded values for the cells the tie-class kernel hands to the general kernel.  Same contract as aucell_csr_f32
 * otherwise; output bit-identical to it on the decoded matrix. */
int aucell_csr16_lut8_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const uint16_t* d_indices,
                          const uint8_t* d_codes, const float* d_lut, float* d_vals_scratch, int64_t n_cells,
                          double* d_out_auc, int32_t* d_out_nnz, void* stream);

/* The same with int32 column indices: 5 bytes per stored entry (a host that prepares the upload only has to touch the
 * values; any n_genes). */
int aucell_csr_lut8_f32(const aucell_plan_t* plan, const int64_t* d_indptr, const int32_t* d_indices,
                        const uint8_t* d_codes, const float* d_lut, float* d_vals_scratch, int64_t n_cells,
                        double* d_out_auc, int32_t* d_out_nnz, void* stream);

/* ---- aucell4r (aucell.py:98-182): AUCs from an existing uint32 ranking matrix [n_cells x n_genes]
 * (row_stride in elements).  The plan's regulon columns index the ranking matrix' columns directly (create the
 * plan with h_perm == NULL). */
int aucell_from_rankings_u32(const aucell_plan_t* plan, const uint32_t* d_rank, int64_t n_cells,
                             int64_t row_stride, double* d_out_auc, void* stream);

/* ---- host-buffer entry points: what a CPU caller (the reference's aucell()) binds -----------------------
 * Inputs and output live in HOST memory (pinned or pageable).  The library streams cell chunks through the
 * device on its own streams (H2D, kernel, D2H overlapped) and returns when h_out_auc is complete.
 * chunk_cells <= 0 picks a default.  h_out_nnz nullable. */
int aucell_csr_f32_host(const aucell_plan_t* plan, const int64_t* h_indptr, const int32_t* h_indices,
                        const float* h_data, int64_t n_cells, double* h_out_auc, int32_t* h_out_nnz,
                        int64_t chunk_cells);
int aucell_dense_f32_host(const aucell_plan_t* plan, const float* h_expr, int64_t n_cells, int64_t row_stride,
                          double* h_out_auc, int32_t* h_out_nnz, int64_t chunk_cells);
/* (replaces the DataFrame form of pyscenic.aucell.aucell, reference src/pyscenic/aucell.py:185-213 -- `exp_mtx` is
 * `pd.DataFrame` (n_cells x n_genes) there; this takes `exp_mtx.values` without a transposing or casting copy.)
 * A host matrix in dense storage as a pandas DataFrame holds it: float32 (is_f64 = 0) or float64 values, cell-major
 * (gene_stride == 1, cell_stride >= n_genes) or gene-major (cell_stride == 1, gene_stride >= n_cells; strides in
 * elements).  Expression matrices are mostly zeros: host threads squeeze them out in one pass at memory speed and the
 * call continues as aucell_csr_f32_host -- a fraction of the bytes over PCIe, and the CSR kernels on the device.
 * Returns AUCELL_NOT_SPARSE (1), with nothing computed, when more than 30 % of a sample of the entries are non-zero,
 * when a float64 value does not survive the cast to float32 (a cast that changes values may create ties the input does
 * not have), or when the plan has no tie-class kernel: the caller then uses aucell_dense_f32_host / device chunks. */
int aucell_dense_strided_host(const aucell_plan_t* plan, const void* h_expr, int is_f64, int64_t n_cells,
                              int64_t cell_stride, int64_t gene_stride, double* h_out_auc, int32_t* h_out_nnz);
/* The host-side half of aucell_dense_strided_host on its own (no GPU): a dense host matrix (float32 / float64,
 * cell-major or gene-major, strides in elements) -> CSR with float32 values and sorted int32 columns, zeros (+0.0 / -0.0)
 * dropped, NaN kept; one pass on all host cores.  h_indptr [n_cells + 1] and *out_nnz are always filled; h_indices /
 * h_values hold `capacity` entries (AUCELL_ERR_NOMEM when that is fewer than *out_nnz: call again with larger arrays).
 * Returns AUCELL_NOT_SPARSE (1), nothing else written, when a float64 value does not survive the cast to float32. */
int aucell_host_dense_to_csr(const void* h_expr, int is_f64, int64_t n_cells, int64_t n_genes, int64_t cell_stride,
                             int64_t gene_stride, int64_t* h_indptr, int32_t* h_indices, float* h_values, int64_t capacity,
                             int64_t* out_nnz);
/* float64 -> float32 on all host cores, with the check the ranking needs: returns 1 when every value survives the cast
 * unchanged (NaN counts as unchanged), 0 when one does not (a cast that changes values may create ties the input does
 * not have: the caller then keeps float64 and uses the *_f64 entry points), negative on bad arguments.  numpy needs
 * three single-threaded passes for the same answer (seconds on a few hundred million stored values). */
int aucell_cast_f64_to_f32(const double* h_src, float* h_dst, int64_t n);
/* int64 -> int32 column indices on all host cores (scipy switches to int64 indices beyond 2^31 stored entries: the
 * 1.3 M-cell matrix): returns 1 when every value lies in [0, 2^31), 0 when not, negative on bad arguments. */
int aucell_cast_i64_to_i32(const int64_t* h_src, int32_t* h_dst, int64_t n);
/* The host-buffer calls keep their staging (four chunk slots per device: a stream, a device buffer, page-locked host
 * buffers -- a few hundred MB) for the life of the process, because a plan is made per aucell() call and page-locking
 * costs more than ranking 300,000 cells.  This frees it (device < 0: on every device). */
int aucell_b200_release_staging(int device);
/* Host-side check of what the CSR entry points require: 1 when every row's column indices are strictly increasing and
 * inside [0, n_genes) (scipy's "canonical format": sorted, no duplicates), 0 when not, negative on bad arguments.  On
 * host threads, at memory speed (scipy's own check is single-threaded); no GPU involved. */
int aucell_csr_is_canonical(const int64_t* h_indptr, const int32_t* h_indices, int64_t n_cells, int64_t n_genes);

/* ---- downstream of the AUC matrix ------------------------------------------------------------------- */
/* Regulon specificity scores (replaces pyscenic.rss.regulon_specificity_scores, reference src/pyscenic/rss.py:8-33):
 * d_out[t * n_regulons + r] = 1 - sqrt(JSD(auc[:, r] / sum(auc[:, r]), [label == t] / count_t)), natural logarithm.
 * d_auc: device fp64 [n_cells x row_stride] (the layout aucell_* writes); d_labels: device int32 type index per cell
 * in [0, n_types); h_type_counts: HOST int64[n_types], cells per type (each > 0, summing to n_cells); d_out: device
 * fp64 [n_types x n_regulons].  Asynchronous on `stream` apart from one small upload; deterministic (no atomics).
 * A regulon whose AUCs are all zero yields NaN, as in the reference. */
int aucell_rss_f64(int device, const double* d_auc, int64_t n_cells, int64_t row_stride, int32_t n_regulons,
                   const int32_t* d_labels, int32_t n_types, const int64_t* h_type_counts, double* d_out, void* stream);

/* Binarization of an AUC matrix (replaces pyscenic.binarization.derive_threshold / binarize, reference
 * src/pyscenic/binarization.py:12-98), all regulons at once.  d_cols: device fp64 [n_regulons x n_cells], row r = the AUC
 * values of regulon r (the transposed AUC matrix); "sorted" entry points need every row in ascending order.  Asynchronous
 * on `stream`, deterministic (no atomics).
 *   stats   d_out4[r]  = { mean, variance (ddof 0), variance (ddof 1), 0 }                      -> mean + 2 std
 *   dip     d_out_dip[r] = Hartigan's dip statistic of row r; d_scratch: int32 [n_regulons x 4 x (n_cells + 1)]
 *   gmm2    two-component Gaussian mixture by EM with sklearn's update rules, started from the hard assignment
 *           "the first d_split[r] sorted values are component 0"; d_out9[r] = { w0, w1, mean0, mean1, var0, var1,
 *           mean log-likelihood under the final parameters, iterations, converged }
 *   trough  scipy's bounded Brent minimiser of the Gaussian kernel density of row r (kernel variance d_kde_var[r])
 *           between d_lo[r] and d_hi[r]; d_out_x[r] = the minimiser
 */
int aucell_binarize_stats_f64(int device, const double* d_cols, int32_t n_regulons, int64_t n_cells, double* d_out4,
                              void* stream);
int aucell_binarize_dip_f64(int device, const double* d_sorted_cols, int32_t n_regulons, int64_t n_cells,
                            int32_t* d_scratch, double* d_out_dip, void* stream);
int aucell_binarize_gmm2_f64(int device, const double* d_sorted_cols, int32_t n_regulons, int64_t n_cells,
                             const int32_t* d_split, int32_t max_iter, double tol, double reg_covar, double* d_out9,
                             void* stream);
int aucell_binarize_trough_f64(int device, const double* d_cols, int32_t n_regulons, int64_t n_cells, const double* d_lo,
                               const double* d_hi, const double* d_kde_var, double xatol, int32_t maxiter,
                               double* d_out_x, void* stream);

/* Non-zero entries per cell of a CSR matrix -- np.count_nonzero(ex_mtx, axis=1), what pyscenic.aucell.derive_auc_threshold
 * takes its quantiles of (reference src/pyscenic/aucell.py:54-72): NaN counts as non-zero, an explicit +-0.0 does not.
 * (The aucell_* entry points deliver the same numbers through d_out_nnz as a by-product.) */
int aucell_count_nonzero_csr_f32(int device, const int64_t* d_indptr, const float* d_data, int64_t n_cells,
                                 int32_t* d_out_nnz, void* stream);

/* ---- instrumentation -------------------------------------------------------------------------------- */
/* Bytes the last aucell_*_host call on this plan moved over PCIe: out2 = { host-to-device, device-to-host }.  (The CSR
 * call uploads uint16 column indices when n_genes <= 65535, so this can be less than the size of the caller's arrays.) */
int aucell_plan_last_host_bytes(const aucell_plan_t* plan, int64_t* out2);

/* Which code path the fused kernel took, summed over every call on this plan since the last reset (synchronises
 * the device): out5 = { cells, general path: more than 4096 explicit entries, general path: negative values ranked
 * below the cutoff, general path: crowded value bin, cells that needed zeros below the cutoff }. */
int aucell_plan_path_stats(const aucell_plan_t* plan, uint64_t* out5, int reset);
/* The CSR entry points with float32 values run two kernels: aucell_tie_kernel ranks every cell whose values fall into
 * a few tie classes (count data), and hands the rest (rows of more than 4096 stored entries, negative / NaN values,
 * continuous values) to aucell_fused_kernel through a to-do list.  Both add the same integers, so the output does not
 * depend on which kernel ranked a cell.  enable = 0 sends every cell to the general kernel, for testing; returns whether
 * the plan qualifies for the tie-class kernel at all (16-bit positions, non-negative weights of bounded range) through
 * *out_available (nullable). */
int aucell_plan_set_fast_path(aucell_plan_t* plan, int enable, int* out_available);
/* Cells per outcome of aucell_tie_kernel since the last reset (synchronises the device): out8 = { ranked there, handed
 * over: more stored entries than a cell group holds, negative / NaN value, duplicate or out-of-range column, two values
 * in one class bin, more than 256 tail entries, skipped (the group saw mostly unsuitable cells), reserved }. */
int aucell_plan_fast_stats(const aucell_plan_t* plan, uint64_t* out8, int reset);
/* number of kernels this library has launched on the calling process since load (bench.py gpu_launches) */
int64_t aucell_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AUCELL_B200_H */
