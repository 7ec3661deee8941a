# R glue for the B200 backend: the ONLY change to scDiffCom's R code (see INTEGRATION.md).
#
# In R/utils_permutation.R, run_stat_analysis() keeps lines 1-115 (row subset, observed vectors, gene subset) and
# 215-499 (p-values, merge, NA / 1 fill, distributions naming) unchanged. The batch loop at lines 116-214 — and the
# single future_replicate() at lines 305-314 when return_distributions is TRUE — is replaced by one call to
# scdc_counts_b200() below when getOption("scDiffCom.backend", Sys.getenv("SCDIFFCOM_BACKEND", "R")) == "b200".
# The backend switch is out-of-band on purpose: validate_parameters() (R/utils_validation.R:34-44) checks the exact
# set of parameter names, so no user-facing argument can be added.
#
# This file has never been executed (R is not installed in the build image); it documents the marshaling contract of
# src/scdc_rcall.c and is kept next to it.

scdc_counts_b200 <- function(analysis_inputs, sub_cci_template_iter, observed, iterations, score_type,
                             return_distributions, seed = 42L, n_gpus = 0L, precision = "fp64", rel_tol = 0) {
  data_tr <- analysis_inputs$data_tr                       # cells x genes dgCMatrix, already gene-subset (:79-84)
  cell_types <- analysis_inputs$cell_types
  n_groups <- length(cell_types) * (if (analysis_inputs$condition$is_cond) 2L else 1L)
  if (n_groups > 254L) {                                   # 8-bit group codes on the device (include/scdc.h)
    stop("scDiffCom B200 backend: more than 254 (cell type x condition) groups; use the R backend")
  }
  ct <- match(analysis_inputs$metadata$cell_type, cell_types) - 1L
  cond <- NULL
  if (analysis_inputs$condition$is_cond) {
    cond <- as.integer(analysis_inputs$metadata$condition == analysis_inputs$condition$cond2)
  }
  LR_COLNAMES <- c(paste0("LIGAND_", 1:analysis_inputs$max_nL), paste0("RECEPTOR_", 1:analysis_inputs$max_nR))
  LRI <- analysis_inputs$LRI
  n_lri <- nrow(LRI)
  sub_gene <- matrix(                                      # (max_nL + max_nR) x n_LRI, 0-based, NA -> -1
    unlist(lapply(LR_COLNAMES, function(col) {             # (explicit shape: sapply() + t() collapses when n_LRI == 1)
      g <- match(LRI[[col]], colnames(data_tr)) - 1L
      g[is.na(g)] <- -1L
      as.integer(g)
    })),
    nrow = length(LR_COLNAMES), ncol = n_lri, byrow = TRUE
  )
  storage.mode(sub_gene) <- "integer"
  rows <- matrix(
    c(match(sub_cci_template_iter$LRI, LRI$LRI) - 1L,
      match(sub_cci_template_iter$EMITTER_CELLTYPE, cell_types) - 1L,
      match(sub_cci_template_iter$RECEIVER_CELLTYPE, cell_types) - 1L),
    ncol = 3
  )
  storage.mode(rows) <- "integer"
  observed <- as.matrix(observed)
  storage.mode(observed) <- "double"
  stopifnot(is.numeric(iterations), length(iterations) == 1L, !is.na(iterations), iterations >= 1,
            is.numeric(seed), length(seed) == 1L, !is.na(seed), seed >= 0)
  opts <- list(
    iterations = as.numeric(iterations), seed = as.numeric(seed), n_gpus = as.integer(n_gpus),
    precision = if (identical(precision, "fp32_fast")) 1L else 0L, rel_tol = as.numeric(rel_tol),
    score_type = if (score_type == "geometric_mean") 0L else 1L,
    n_celltypes = length(cell_types), max_nL = analysis_inputs$max_nL, max_nR = analysis_inputs$max_nR,
    return_distributions = as.integer(return_distributions),
    perm_cond = NULL, perm_ct = NULL                       # raw matrices N x iterations for validation mode
  )
  .Call(
    "C_scdc_perm_counts",
    as.integer(data_tr@p), as.integer(data_tr@i), as.numeric(data_tr@x), as.integer(data_tr@Dim),
    as.integer(ct), if (is.null(cond)) NULL else as.integer(cond), sub_gene, rows, observed, opts,
    PACKAGE = "scDiffCom"
  )
}

# Inside run_stat_analysis(), replacing R/utils_permutation.R:116-214 (counting branch):
#
#   if (identical(getOption("scDiffCom.backend", Sys.getenv("SCDIFFCOM_BACKEND", "R")), "b200")) {
#     observed <- if (!analysis_inputs$condition$is_cond) {
#       matrix(cci_score_actual, ncol = 1)
#     } else {
#       do.call(cbind, c(list(cci_score_diff_actual, cci_score_cond1_actual, cci_score_cond2_actual),
#                        ligand_score_diff_actual, receptor_score_diff_actual))
#     }
#     n_run <- if (iterations <= internal_iter) iterations else
#       (n_broad_iter - 1) * internal_iter + ifelse(iterations %% internal_iter == 0, internal_iter, iterations %% internal_iter)
#     array_counts <- scdc_counts_b200(analysis_inputs, sub_cci_template_iter, observed, n_run, score_type, FALSE,
#                                      seed = <seed of run_interaction_analysis>)
#     n_broad_iter <- 1          # array_counts is already summed over batches: lines 215-259 then divide by `iterations`
#     if (analysis_inputs$condition$is_cond) {
#       array_counts <- array(as.vector(array_counts), dim = c(nrow(array_counts), ncol(array_counts), 1))
#     } else {
#       array_counts <- as.vector(array_counts)   # plain vector (no dim, no "stats" attribute): `P_VALUE := pvals` at :215-221
#     }
#   } else { <original sapply(...) loop> }
# getOption("scDiffCom.b200.precision", "fp64") / getOption("scDiffCom.b200.rel_tol", 0) select SCDC_FP32_FAST the same
# out-of-band way; the band counters come back as attr(array_counts, "band_counts") (read it before as.vector()).
