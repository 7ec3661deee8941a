/*
 * scdc.h — C-ABI of the B200-native permutation-test engine behind scDiffCom's run_stat_analysis().
 *
 * This is the drop-in boundary. The reference has NO native interface today; the only
 * "operator" on this path is the internal R function
 *     run_stat_analysis(analysis_inputs, cci_dt_simple, iterations, return_distributions, score_type, verbose)
 *     (reference R/utils_permutation.R:1-8, returns list(cci_raw, distributions) at :470-498).
 * Its batch loop (R/utils_permutation.R:116-214, and :305-314 for return_distributions) together with
 * run_stat_iteration (R/utils_permutation.R:501-603) and the compute_fast branch of run_simple_cci_analysis
 * (R/utils_cci.R:171-261: aggregate_cells :441-462 + build_cci_or_drate("cci") :464-712) is what
 * scdc_perm_counts() replaces. Everything before (:1-115: row subset, observed vectors, gene subset) and after
 * (:215-499: p-values = counts / iterations, merge, NA / 1 fill) stays in R. The `.Call` wrapper an R maintainer
 * adds (src/scdc_rcall.c) and the R glue are shown in INTEGRATION.md.
 *
 * Plain C: pointers and sizes only, no R types, no torch types, no C++ in the signatures. The library never
 * throws, never calls exit/abort, never retains caller memory past return. There is NO CPU fallback: without a
 * visible sm_100 device every compute entry point returns SCDC_ENOGPU.
 */
#ifndef SCDC_H
#define SCDC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCDC_VERSION_STRING "0.2.0"

/* error codes (0 = success) */
enum {
  SCDC_OK = 0,
  SCDC_EINVAL = 1, /* bad argument / inconsistent problem (message in scdc_last_error) */
  SCDC_ENOGPU = 2, /* no visible CUDA device of compute capability 10.x */
  SCDC_ECUDA = 3,  /* CUDA runtime failure */
  SCDC_ENCCL = 4,  /* NCCL failure (multi-GPU merge of the counters) */
  SCDC_ENOMEM = 5,  /* host or device allocation failure */
  SCDC_EINTERRUPT = 6 /* scdc_options.interrupt asked to stop; the counters of this call are discarded */
};

/* score_type (reference argument `score_type`, R/utils_cci.R:605,659) */
enum { SCDC_SCORE_GEOMETRIC = 0, SCDC_SCORE_ARITHMETIC = 1 };

/* precision modes.
 *   SCDC_FP64_EXACT: fp64 values and sums in the reference's summation order; counts are bit-identical to the reference
 *                    arithmetic for the same labels.
 *   SCDC_FP32_FAST : values and group sums in fp32 (same summation order), scores / differences / comparisons in fp64
 *                    from the fp32 means. Counts may differ from the exact ones only for comparisons that fall inside
 *                    the tolerance band (scdc_options.rel_tol); every such comparison is reported in
 *                    scdc_result.band_counts, so |count_fp32 - count_fp64| <= band_count per (row, column). */
enum { SCDC_FP64_EXACT = 0, SCDC_FP32_FAST = 1 };

/* counts[] entry for a (row, column) whose observed value is NA (absent subunit): the reference yields NA there
 * (R/utils_permutation.R:442-469); the .Call wrapper maps this to NA_real_. */
#define SCDC_COUNT_NA UINT64_MAX

/*
 * The problem = analysis_inputs (R/utils_preprocessing.R:41-49, after the gene subset of
 * R/utils_permutation.R:79-84) + sub_cci_template_iter (R/utils_permutation.R:85,114) + the observed vectors
 * (R/utils_permutation.R:13,29-63), integer-encoded once by the caller:
 *   ct_code   = match(metadata$cell_type, cell_types) - 1
 *   cond_code = 0 for cond1, 1 for cond2 (NULL in detection mode)
 *   sub_gene  = match(LIGAND_i / RECEPTOR_j, colnames(data_tr)) - 1, NA -> -1; one row per kept LRI,
 *               row-major [n_lri][max_nL + max_nR] in the order L_1..L_nL, R_1..R_nR
 *   row_*     = per tested row: index into the LRI table, emitter and receiver cell-type codes
 *   observed  = column-major [n_rows x ncols]; ncols = 1 (detection: CCI_SCORE) or 3 + max_nL + max_nR
 *               (differential: [DE = cond2 - cond1, cond1, cond2, L_1..L_nL diffs, R_1..R_nR diffs],
 *               the column order of R/utils_permutation.R:580-598). NaN marks NA.
 * The expression matrix is data_tr as it is: cells x genes dgCMatrix slots p, i, x (0-based CSC; within a gene
 * column the cell indices ascend).
 */
typedef struct scdc_problem {
  int32_t n_cells;      /* N */
  int32_t n_genes;      /* G */
  int32_t n_celltypes;  /* T, with T * n_conditions <= 254 (group codes are 8-bit on the device): T <= 127 in differential
                           mode, T <= 254 in detection mode; larger inputs are rejected with SCDC_EINVAL and the R glue
                           keeps the reference's R loop for them */
  int32_t n_conditions; /* C: 1 = detection mode, 2 = differential mode */
  int32_t n_lri;        /* L */
  int32_t n_rows;       /* R_sub */
  int32_t max_nL;       /* 1..2 (any >= 1 accepted) */
  int32_t max_nR;       /* 1..3 (any >= 1 accepted) */
  const int32_t* colptr;   /* [G + 1] */
  const int32_t* rowidx;   /* [nnz] */
  const double* values;    /* [nnz] */
  const int32_t* ct_code;  /* [N] */
  const int32_t* cond_code;/* [N] or NULL */
  const int32_t* sub_gene; /* [L * (max_nL + max_nR)] */
  const int32_t* row_lri;  /* [R_sub] */
  const int32_t* row_emit; /* [R_sub] */
  const int32_t* row_recv; /* [R_sub] */
  const double* observed;  /* [R_sub * ncols], column-major */
} scdc_problem;

typedef struct scdc_options {
  int64_t iterations;   /* replicates to run in THIS call (>= 1). p = counts / iterations is left to the caller */
  uint64_t seed;        /* Philox key (reference: set.seed(seed), R/interaction_analysis.R:245) */
  int64_t perm_offset;  /* global index of this call's first replicate: Philox counter = perm_offset + i, so any
                           partition of [0, total) over calls / GPUs / processes reproduces the same labels */
  int32_t score_type;   /* SCDC_SCORE_* */
  int32_t n_gpus;       /* scdc_perm_counts only: 0 = all visible devices, k = first k devices */
  int32_t precision;    /* SCDC_FP64_EXACT (default) or SCDC_FP32_FAST */
  int32_t batch_size;   /* replicates per device launch; 0 = auto */
  /* validation mode: the reference's own permuted labels, row-major [iterations][N] (replicate-major).
     detection: perm_ct = sample(cell_type) codes (R/utils_permutation.R:509); perm_cond must be NULL.
     differential: perm_cond = conditions after :536-541, perm_ct = cell types after :552-569. Both NULL -> the
     labels are drawn on the device (Philox4x32-10, same strata, same distribution as base::sample). */
  const uint8_t* perm_cond;
  const uint8_t* perm_ct;
  int32_t return_distributions; /* 1: also fill result.distributions (R/utils_permutation.R:304-426) */
  int32_t device_base;  /* scdc_perm_counts only: index (among the visible compute-capability-10.x devices) of the
                           first device to use; 0 = default. One process per GPU sets it to its local rank. */
  void* stream;         /* scdc_run only: NULL = the handle's own stream, else a cudaStream_t of the handle's device on
                           which all work of the call is enqueued (lets a caller time it with its own CUDA events) */
  /* Optional. Polled on the CALLING thread between device batches (never from the library's own threads); a non-zero
     return abandons the call with SCDC_EINTERRUPT. The .Call wrapper passes a function that runs
     R_CheckUserInterrupt() under R_ToplevelExec (src/scdc_rcall.c), which is how a long permutation run stays
     interruptible from the R console, as the reference's future_replicate loop is (R/utils_permutation.R:129-138). */
  int (*interrupt)(void* user);
  void* interrupt_user;
  /* SCDC_FP32_FAST only: relative half-width of the tie band (<= 0: the default 1e-5). A comparison `perm >= obs` is
     "in the band" when |perm - obs| <= rel_tol * scale, scale = |obs| for the one-sided score columns (geometric scores
     are compared as products: |x - obs^2| <= 2 rel_tol obs^2) and = the sum of the two magnitudes being subtracted for
     the two-sided difference columns (DE: score_cond2 + score_cond1; subunits: mean_cond2 + mean_cond1), which is the
     scale of the fp32 rounding error of a difference. */
  double rel_tol;
} scdc_options;

typedef struct scdc_stats {
  double ms_total;      /* host wall time inside the call */
  double ms_upload;     /* H2D of the problem (0 when a handle is reused) */
  double ms_labels;     /* device time, label generation / upload + transpose (CUDA events) */
  double ms_reduce;     /* device time, group-sum kernels (CUDA events) */
  double ms_score;      /* device time, score / compare / count kernels (CUDA events) */
  double ms_merge;      /* counter merge across GPUs + D2H */
  double ms_device;     /* device time of the whole batch loop, first to last CUDA event (max over GPUs) */
  double reduce_bytes;  /* algorithmic bytes of all group-sum launches ("algorithmic bytes") */
  double reduce_wavefronts; /* shared-memory wavefronts the group-sum kernels must issue by design: the unit
                           that actually bounds them (one wavefront per clock per SM) */
  int64_t reduce_launches;
  int64_t kernel_launches; /* all kernels launched by the call */
  int64_t batch_size;   /* replicates per launch actually used */
  int32_t n_gpus;       /* devices actually used */
  int32_t reserved;     /* multi-GPU calls: 1 = NCCL merged the counters, 0 = peer copies + adds (NCCL unavailable) */
  double h2d_bytes;     /* bytes copied host -> device by the call (problem, rows, uploaded labels) */
  double d2h_bytes;     /* bytes copied device -> host (counters, distributions) */
  double p2p_bytes;     /* bytes copied device -> device over NVLink (multi-GPU: the problem is uploaded once and broadcast) */
} scdc_stats;

typedef struct scdc_result {
  uint64_t* counts;       /* [R_sub * ncols] column-major, caller-allocated; SCDC_COUNT_NA where observed is NaN */
  double* distributions;  /* NULL, or caller-allocated [R_sub * ncols_d * iterations] (replicate-major blocks of a
                             column-major R_sub x ncols_d matrix); ncols_d = 1 (detection) or 3 (DE, cond1, cond2) */
  void* counts_device;    /* scdc_run only: NULL, or a DEVICE buffer of int64[R_sub * ncols] on the handle's device
                             that receives the raw counts (NA entries = 0) for a caller-side collective */
  uint64_t* band_counts;  /* NULL, or caller-allocated [R_sub * ncols] column-major: comparisons inside the tolerance band
                             (SCDC_FP32_FAST; all zero in SCDC_FP64_EXACT, where nothing is approximated). NA entries as
                             in counts */
  void* band_device;      /* scdc_run only: like counts_device, for the band counters */
  scdc_stats stats;
} scdc_result;

/* One-shot entry point: what `.Call("C_scdc_perm_counts", ...)` binds. Uploads the problem, runs `iterations`
 * replicates sharded over `n_gpus` devices (contiguous replicate ranges, replicated expression matrix), merges the
 * integer counters with one ncclAllReduce(sum) and writes them to result->counts. */
int scdc_perm_counts(const scdc_problem* problem, const scdc_options* options, scdc_result* result);

/* Handle API: keep the problem resident on one device across calls (benchmarks, resumable / sharded runs:
 * counts are additive over replicate ranges). */
typedef struct scdc_handle scdc_handle;
int scdc_create(const scdc_problem* problem, int32_t device, scdc_handle** out);
int scdc_run(scdc_handle* handle, const scdc_options* options, scdc_result* result);
void scdc_release(scdc_handle* handle);

/* Observed-pass helper (reference aggregate_cells, R/utils_cci.R:441-462, on the ORIGINAL labels): group means of the
 * expression and of 1*(x > 0) (detection rate, R/utils_cci.R:262-266), each [K x G] column-major with K = C * T and
 * group index = cond_code * T + ct_code. Either output may be NULL. */
int scdc_group_means(scdc_handle* handle, double* means, double* detection_rate);

/* (Re)install the tested rows on a handle: sub_cci_template_iter + observed vectors (reference
 * R/utils_permutation.R:11-64). Lets one resident handle serve the observed pass first (create it with n_rows = 0), then
 * the permutation test; only the genes the rows reference are reduced afterwards (the reference's gene subset, :79-84). */
int scdc_set_rows(scdc_handle* handle, int32_t n_rows, const int32_t* row_lri, const int32_t* row_emit,
                  const int32_t* row_recv, const double* observed);

/* Observed pass over n template rows (reference run_simple_cci_analysis(compute_fast = FALSE), R/utils_cci.R:171-283,
 * SURVEY.md "next" item f2): every output is caller-allocated, column-major over the n rows, and may be NULL.
 *   expression / detection_rate [n x (C * nS)]: column c * nS + s = subunit s (L_1..L_nL, R_1..R_nR) in condition c, taken in
 *       the emitter (ligands) / receiver (receptors); NaN = NA subunit          (L{i}_EXPRESSION*, R{j}_DETECTION_RATE* ...)
 *   score [n x C]: CCI_SCORE per condition;  is_expressed [n x C]: IS_CCI_EXPRESSED per condition (is_detected_full with
 *       threshold_pct / threshold_min_cells, R/utils_cci.R:783-802).
 * LOGFC (a log of a score ratio) is left to the caller. */
typedef struct scdc_observed {
  double* expression;
  double* detection_rate;
  double* score;
  uint8_t* is_expressed;
} scdc_observed;
int scdc_observed_pass(scdc_handle* handle, int64_t n, const int32_t* row_lri, const int32_t* row_emit,
                       const int32_t* row_recv, double threshold_pct, double threshold_min_cells, int32_t score_type,
                       scdc_observed* out);

/* ---- "next" item f3: the CJ template and the merge-back, integer-coded, from the shim --------------
 * Host-side helpers (no device needed): at T = 60 the reference builds these with CJ() + string-keyed merges over 18 M rows
 * (R/utils_cci.R:1-169) and merges the p-values back with another keyed merge (R/utils_permutation.R:427-469); here they are
 * O(rows) integer loops over a few host threads, and R attaches the factor levels once.
 *
 * scdc_cci_template = create_cci_template + add_cell_number (R/utils_cci.R:1-30, 89-169). Row i of the template
 * (CJ order: emitter, receiver, LRI name ascending) is emitter = i / (T * L), receiver = (i / L) % T, LRI = lri_order[i % L],
 * where lri_order[L] lists the LRI indices sorted by name (order(LRI$LRI) - 1L; CJ sorts its arguments). Outputs, each
 * caller-allocated with n = T * T * L entries (ncells_*: [C][n], condition-major): emitter / receiver cell-type codes, LRI
 * index, and NCELLS_EMITTER / NCELLS_RECEIVER (per condition in differential mode). Any output may be NULL. */
int scdc_cci_template(int32_t n_celltypes, int32_t n_conditions, int32_t n_lri, const int32_t* lri_order, int32_t n_cells,
                      const int32_t* ct_code, const int32_t* cond_code, int32_t* emitter, int32_t* receiver,
                      int32_t* lri_idx, int32_t* ncells_emitter, int32_t* ncells_receiver);

/* scdc_merge_pvalues = R/utils_permutation.R:215-259 (p = counts / iterations) + :427-469 (merge back into the full table):
 * pvalues is caller-allocated, column-major [n_full x ncols]; rows outside the tested subset get 1 in every column
 * (:434-441); tested row k (position tested_rows[k] of the full table) gets counts[k] / iterations, NaN for SCDC_COUNT_NA;
 * and in differential mode (ncols = 3 + max_nL + max_nR) the subunit columns are NaN (= NA) wherever the row's LRI lacks
 * that subunit (:442-469), tested or not. lri_idx[n_full] = LRI index of every full row; sub_gene as in scdc_problem. */
int scdc_merge_pvalues(int64_t n_full, int64_t n_tested, const int64_t* tested_rows, int32_t ncols, const uint64_t* counts,
                       int64_t iterations, int32_t n_lri, int32_t max_nL, int32_t max_nR, const int32_t* sub_gene,
                       const int32_t* lri_idx, double* pvalues);

/* ---- "next" item f4 (optional polish; north_star keeps filtering in R): stats::p.adjust(p, method = "BH") as the reference
 * applies it to the permutation p-values in process_cci_raw (R/utils_filtering.R:216-235): NaN entries (NA) are left out
 * of the ranking and stay NaN; out[i] = min(1, min over j with p_j >= p_i of n p_j / rank_j), n = number of non-NaN
 * entries. `group` (may be NULL) = one integer key per entry: the adjustment is done separately inside every group (the
 * reference's `by = temp_by`). Host-side helper; out may alias p. Counts themselves are resumable without any extra entry
 * point: they are additive over replicate ranges (scdc_options.perm_offset), so a caller may add the counts of successive
 * calls and divide by the total. */
int scdc_bh_adjust(int64_t n, const double* p, const int32_t* group, double* out);

/* Debug / test hook: the labels scdc_run would draw on the device for replicates [perm_offset, perm_offset + n),
 * row-major [n][N]. cond_out may be NULL in detection mode. */
int scdc_draw_labels(scdc_handle* handle, uint64_t seed, int64_t perm_offset, int64_t n, uint8_t* cond_out,
                     uint8_t* ct_out);

/* Debug / test hook: checks the DE-column decision procedure of the score kernel (rsqrt.approx band filter + IEEE sqrt
 * fall-back) against the plain IEEE expression |sqrt(x2) - sqrt(x1)| >= |observed| on n adversarial inputs.
 * out[0] = mismatches (must be 0), out[1] = decisions taken by the fast path, out[2] = approximations outside 2^-20. */
int scdc_selftest_de_filter(uint64_t n, uint64_t seed, uint64_t out[3]);

/* Device work buffers are drawn from the device's default stream-ordered memory pool and kept there between calls
 * (cudaMallocAsync with the release threshold lifted), so repeated calls do not pay cudaMalloc / cudaFree again.
 * scdc_trim(device) hands the cached memory back to the driver (device = index among all CUDA devices, -1 = all). */
int scdc_trim(int32_t device);

const char* scdc_last_error(void); /* thread-local message of the last failing call on this thread */
int scdc_device_count(void);       /* visible devices with compute capability 10.x (0 when there is none) */
const char* scdc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCDC_H */
