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
                             return_distributions, seed = 42L, n_gpus = 0L) {
  data_tr <- analysis_inputs$data_tr                       # cells x genes dgCMatrix, already gene-subset (:79-84)
  cell_types <- analysis_inputs$cell_types
  ct <- match(analysis_inputs$metadata$cell_type, cell_types) - 1L
  cond <- NULL
  if (analysis_inputs$condition$is_cond) {
    cond <- as.integer(analysis_inputs$metadata$condition == analysis_inputs$condition$cond2)
  }
  LR_COLNAMES <- c(paste0("LIGAND_", 1:analysis_inputs$max_nL), paste0("RECEPTOR_", 1:analysis_inputs$max_nR))
  LRI <- analysis_inputs$LRI
  sub_gene <- t(sapply(LR_COLNAMES, function(col) {
    g <- match(LRI[[col]], colnames(data_tr)) - 1L
    g[is.na(g)] <- -1L
    g
  }))                                                      # (max_nL + max_nR) x n_LRI, 0-based, NA -> -1
  storage.mode(sub_gene) <- "integer"
  rows <- cbind(
    match(sub_cci_template_iter$LRI, LRI$LRI) - 1L,
    match(sub_cci_template_iter$EMITTER_CELLTYPE, cell_types) - 1L,
    match(sub_cci_template_iter$RECEIVER_CELLTYPE, cell_types) - 1L
  )
  storage.mode(rows) <- "integer"
  opts <- list(
    iterations = iterations, seed = seed, n_gpus = n_gpus,
    score_type = if (score_type == "geometric_mean") 0L else 1L,
    n_celltypes = length(cell_types), max_nL = analysis_inputs$max_nL, max_nR = analysis_inputs$max_nR,
    return_distributions = as.integer(return_distributions),
    perm_cond = NULL, perm_ct = NULL                       # raw matrices N x iterations for validation mode
  )
  .Call(
    "C_scdc_perm_counts",
    data_tr@p, data_tr@i, data_tr@x, data_tr@Dim, ct, cond, sub_gene, rows, observed, opts,
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
#     if (analysis_inputs$condition$is_cond) dim(array_counts) <- c(nrow(array_counts), ncol(array_counts), 1)
#   } else { <original sapply(...) loop> }
