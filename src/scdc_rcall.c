/*
 * scdc_rcall.c — the `.Call` wrapper an scDiffCom maintainer adds under src/ (the ONLY file that includes Rinternals.h).
 *
 * Marshals R vectors to the plain-C ABI of include/scdc.h and back. It replaces the batch loop of run_stat_analysis()
 * (reference R/utils_permutation.R:116-214 and :305-314); see INTEGRATION.md for the R glue that calls it and for
 * NAMESPACE / DESCRIPTION / Makevars. R owns every buffer (allocated with Rf_alloc*, PROTECTed here); the engine copies
 * what it needs to the device and keeps nothing past return. Errors are reported with Rf_error() only after the engine
 * call has returned (no C++ frames are alive in this translation unit; Rf_error longjmps).
 *
 * R is not installed in the build image of this repo, so this file is compile-checked against the minimal stub headers
 * in tests/r_stub/ (tests/test_abi.py); it has never run under a real R.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "scdc.h"

static SEXP list_get(SEXP list, const char* name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (int i = 0; i < Rf_length(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

static double opt_num(SEXP opts, const char* name, double dflt) {
  SEXP v = list_get(opts, name);
  return Rf_isNull(v) ? dflt : Rf_asReal(v);
}

/* a finite number in [lo, hi], else an R error naming the option (NA, NaN, negative seeds ... never reach a C cast) */
static double opt_checked(SEXP opts, const char* name, double dflt, double lo, double hi) {
  const double v = opt_num(opts, name, dflt);
  if (ISNAN(v) || v < lo || v > hi) Rf_error("C_scdc_perm_counts: option '%s' must be a number in [%g, %g]", name, lo, hi);
  return v;
}

/*
 * .Call("C_scdc_perm_counts", p, i, x, dim, ct, cond, sub_gene, rows, observed, opts)
 *   p, i, x   data_tr@p, data_tr@i, data_tr@x of the cells x genes dgCMatrix (after R/utils_permutation.R:79-84)
 *   dim       data_tr@Dim = c(N, G)
 *   ct        integer(N)  match(metadata$cell_type, cell_types) - 1L
 *   cond      integer(N)  0L = cond1, 1L = cond2, or NULL in detection mode
 *   sub_gene  integer matrix with (max_nL + max_nR) ROWS and one COLUMN per kept LRI (0-based gene index, NA -> -1L)
 *   rows      integer matrix R_sub x 3: LRI index (0-based column of sub_gene), emitter code, receiver code
 *   observed  double matrix R_sub x ncols in the column order of R/utils_permutation.R:580-598 (NA allowed)
 *   opts      list(iterations, seed, score_type = 0L|1L, n_gpus, n_celltypes, max_nL, max_nR, return_distributions,
 *                  precision = 0L (fp64-exact) | 1L (fp32-fast), rel_tol,
 *                  perm_cond = raw matrix N x iterations | NULL, perm_ct = raw matrix N x iterations | NULL)
 * Returns a double matrix R_sub x ncols of exceedance counts (NA_real_ where the subunit is NA), with attributes
 * "stats" (named numeric), "band_counts" (fp32-fast: double matrix R_sub x ncols, comparisons inside the tolerance band)
 * and, when requested, "distributions" (double array R_sub x ncols_d x iterations).
 */
/* scdc_options.interrupt: R_CheckUserInterrupt() longjmps when the user pressed Ctrl-C / Esc, which must not unwind through
 * the engine's C++ frames; under R_ToplevelExec the jump is caught and reported as FALSE instead. The engine then stops
 * enqueuing work, returns SCDC_EINTERRUPT, and the wrapper raises the R-level condition after the call has returned. */
static void check_interrupt_fn(void* unused) { (void)unused; R_CheckUserInterrupt(); }
static int poll_user_interrupt(void* unused) { (void)unused; return R_ToplevelExec(check_interrupt_fn, NULL) == FALSE; }

SEXP C_scdc_perm_counts(SEXP p, SEXP i, SEXP x, SEXP dim, SEXP ct, SEXP cond, SEXP sub_gene, SEXP rows, SEXP observed,
                        SEXP opts) {
  scdc_problem prob;
  scdc_options opt;
  scdc_result res;
  memset(&prob, 0, sizeof prob);
  memset(&opt, 0, sizeof opt);
  memset(&res, 0, sizeof res);
  if (TYPEOF(p) != INTSXP || TYPEOF(i) != INTSXP || TYPEOF(x) != REALSXP || TYPEOF(ct) != INTSXP ||
      TYPEOF(sub_gene) != INTSXP || TYPEOF(rows) != INTSXP || TYPEOF(observed) != REALSXP || TYPEOF(dim) != INTSXP ||
      (!Rf_isNull(cond) && TYPEOF(cond) != INTSXP) || TYPEOF(opts) != VECSXP)
    Rf_error("C_scdc_perm_counts: wrong argument types");
  if (Rf_length(dim) != 2 || INTEGER(dim)[0] < 1 || INTEGER(dim)[1] < 1) Rf_error("dim must be c(n_cells, n_genes)");
  prob.n_cells = INTEGER(dim)[0];
  prob.n_genes = INTEGER(dim)[1];
  if (Rf_xlength(p) != (R_xlen_t)prob.n_genes + 1) Rf_error("length(p) must be n_genes + 1");
  if (Rf_xlength(i) != Rf_xlength(x) || Rf_xlength(i) < (R_xlen_t)INTEGER(p)[prob.n_genes])
    Rf_error("i and x must have p[n_genes + 1] entries");
  if (Rf_xlength(ct) != prob.n_cells || (!Rf_isNull(cond) && Rf_xlength(cond) != prob.n_cells))
    Rf_error("ct and cond must have one entry per cell");
  if (Rf_ncols(rows) != 3) Rf_error("rows must have 3 columns (LRI index, emitter code, receiver code)");
  prob.n_celltypes = (int32_t)opt_num(opts, "n_celltypes", 0);
  prob.n_conditions = Rf_isNull(cond) ? 1 : 2;
  prob.max_nL = (int32_t)opt_num(opts, "max_nL", 1);
  prob.max_nR = (int32_t)opt_num(opts, "max_nR", 1);
  prob.n_lri = Rf_ncols(sub_gene);
  prob.n_rows = Rf_nrows(rows);
  if (Rf_nrows(sub_gene) != prob.max_nL + prob.max_nR) Rf_error("sub_gene must have max_nL + max_nR rows");
  const int ncols = prob.n_conditions == 2 ? 3 + prob.max_nL + prob.max_nR : 1;
  if (Rf_nrows(observed) != prob.n_rows || Rf_ncols(observed) != ncols) Rf_error("observed has the wrong shape");
  prob.colptr = INTEGER(p);
  prob.rowidx = INTEGER(i);
  prob.values = REAL(x);
  prob.ct_code = INTEGER(ct);
  prob.cond_code = Rf_isNull(cond) ? NULL : INTEGER(cond);
  prob.sub_gene = INTEGER(sub_gene);                 /* column-major (nS x L) == row-major [L][nS] */
  prob.row_lri = INTEGER(rows);
  prob.row_emit = INTEGER(rows) + prob.n_rows;
  prob.row_recv = INTEGER(rows) + 2 * (size_t)prob.n_rows;
  prob.observed = REAL(observed);                    /* NA_real_ is a NaN: the engine treats it as NA */

  opt.iterations = (int64_t)opt_checked(opts, "iterations", 1000, 1, 2e9);
  opt.seed = (uint64_t)opt_checked(opts, "seed", 42, 0, 9007199254740992.0 /* 2^53: exact in a double */);
  opt.score_type = (int32_t)opt_checked(opts, "score_type", SCDC_SCORE_GEOMETRIC, 0, 1);
  opt.n_gpus = (int32_t)opt_checked(opts, "n_gpus", 0, 0, 1024);
  opt.precision = (int32_t)opt_checked(opts, "precision", SCDC_FP64_EXACT, 0, 1);
  opt.rel_tol = opt_checked(opts, "rel_tol", 0, 0, 0.5);
  opt.return_distributions = (int32_t)opt_checked(opts, "return_distributions", 0, 0, 1);
  SEXP pc = list_get(opts, "perm_cond"), pt = list_get(opts, "perm_ct");
  const R_xlen_t n_lab = (R_xlen_t)opt.iterations * prob.n_cells;
  if ((!Rf_isNull(pc) && (TYPEOF(pc) != RAWSXP || Rf_xlength(pc) != n_lab)) ||
      (!Rf_isNull(pt) && (TYPEOF(pt) != RAWSXP || Rf_xlength(pt) != n_lab)))
    Rf_error("perm_cond / perm_ct must be raw matrices n_cells x iterations");
  opt.perm_cond = Rf_isNull(pc) ? NULL : RAW(pc);    /* raw matrix N x iterations, column-major == [iterations][N] */
  opt.perm_ct = Rf_isNull(pt) ? NULL : RAW(pt);
  opt.interrupt = poll_user_interrupt;
  opt.interrupt_user = NULL;

  int nprot = 0;
  const size_t n = (size_t)prob.n_rows * ncols;
  uint64_t* counts = (uint64_t*)R_alloc(n ? n : 1, sizeof(uint64_t));   /* R-managed transient memory */
  res.counts = counts;
  uint64_t* band = NULL;
  if (opt.precision == SCDC_FP32_FAST) {
    band = (uint64_t*)R_alloc(n ? n : 1, sizeof(uint64_t));
    res.band_counts = band;
  }
  SEXP dist = R_NilValue;
  if (opt.return_distributions) {
    const int ncd = prob.n_conditions == 2 ? 3 : 1;
    dist = PROTECT(Rf_alloc3DArray(REALSXP, prob.n_rows, ncd, (int)opt.iterations));
    ++nprot;
    res.distributions = REAL(dist);
  }
  const int rc = scdc_perm_counts(&prob, &opt, &res);
  if (rc != SCDC_OK) {
    UNPROTECT(nprot);
    if (rc == SCDC_EINTERRUPT) Rf_error("scDiffCom B200 backend: interrupted by the user");
    Rf_error("scDiffCom B200 backend: %s (code %d)", scdc_last_error(), rc);
  }
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, prob.n_rows, ncols));
  ++nprot;
  double* o = REAL(out);
  for (size_t k = 0; k < n; ++k) o[k] = (counts[k] == SCDC_COUNT_NA) ? NA_REAL : (double)counts[k];
  static const char* stat_names[] = {"ms_total", "ms_upload", "ms_labels", "ms_reduce", "ms_score", "ms_merge",
                                     "ms_device", "batch_size", "n_gpus", "kernel_launches"};
  const double stat_vals[] = {res.stats.ms_total, res.stats.ms_upload, res.stats.ms_labels, res.stats.ms_reduce,
                              res.stats.ms_score, res.stats.ms_merge, res.stats.ms_device, (double)res.stats.batch_size,
                              (double)res.stats.n_gpus, (double)res.stats.kernel_launches};
  const int ns = (int)(sizeof stat_vals / sizeof stat_vals[0]);
  SEXP stats = PROTECT(Rf_allocVector(REALSXP, ns));
  SEXP snames = PROTECT(Rf_allocVector(STRSXP, ns));
  nprot += 2;
  for (int k = 0; k < ns; ++k) {
    REAL(stats)[k] = stat_vals[k];
    SET_STRING_ELT(snames, k, Rf_mkChar(stat_names[k]));
  }
  Rf_setAttrib(stats, R_NamesSymbol, snames);
  Rf_setAttrib(out, Rf_install("stats"), stats);
  if (band) {
    SEXP bm = PROTECT(Rf_allocMatrix(REALSXP, prob.n_rows, ncols));
    ++nprot;
    for (size_t k = 0; k < n; ++k) REAL(bm)[k] = (band[k] == SCDC_COUNT_NA) ? NA_REAL : (double)band[k];
    Rf_setAttrib(out, Rf_install("band_counts"), bm);
  }
  if (opt.return_distributions) Rf_setAttrib(out, Rf_install("distributions"), dist);
  UNPROTECT(nprot);
  return out;
}

SEXP C_scdc_device_count(void) { return Rf_ScalarInteger(scdc_device_count()); }

static const R_CallMethodDef call_methods[] = {
    {"C_scdc_perm_counts", (DL_FUNC)&C_scdc_perm_counts, 10},
    {"C_scdc_device_count", (DL_FUNC)&C_scdc_device_count, 0},
    {NULL, NULL, 0}};

void R_init_scDiffCom(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
