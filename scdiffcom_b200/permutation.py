"""`run_stat_analysis` — host-side mirror of the reference's permutation driver (R/utils_permutation.R:1-499).

Same name, argument meaning, outputs and error behaviour as the R function. The set-up half (:1-115: tested-row subset,
observed vectors, gene subset) and the tear-down half (:215-499: p = counts / iterations, merge-back, 1 / NA fill,
distribution matrices) are restated here in NumPy; the batch loop in between (:116-214 and :305-314, i.e.
future_replicate(run_stat_iteration)) is ONE call into the B200 engine through the C-ABI (`scdc_perm_counts`).
There is no CPU path for that call.
"""
from __future__ import annotations

import numpy as np

from . import cci as _cci
from .engine import PermutationProblem, merge_pvalues, perm_counts

INTERNAL_ITER = 1000   # R/utils_permutation.R:10


def iterations_to_run(iterations, return_distributions=False):
    """Replicates the reference actually executes (batching quirk, R/utils_permutation.R:100-128): for iterations > 1000
    it runs floor(iterations/1000) batches whose last one has iterations %% 1000 (or 1000) replicates, while p-values
    divide by `iterations`. Identical to `iterations` for iterations <= 1000 or multiples of 1000."""
    if iterations <= INTERNAL_ITER:
        return iterations
    if return_distributions:
        raise ValueError("Only a maximum of 1000 iterations is allowed when 'return_distributions' is TRUE")
    n_broad = iterations // INTERNAL_ITER
    last = iterations % INTERNAL_ITER or INTERNAL_ITER
    return (n_broad - 1) * INTERNAL_ITER + last


def tested_rows(ai, cci_dt_simple):
    """R/utils_permutation.R:11-28."""
    cond = ai["condition"]
    if not cond["is_cond"]:
        return np.asarray(cci_dt_simple["IS_CCI_EXPRESSED"], dtype=bool)
    return (np.asarray(cci_dt_simple[f"IS_CCI_EXPRESSED_{cond['cond1']}"], dtype=bool)
            | np.asarray(cci_dt_simple[f"IS_CCI_EXPRESSED_{cond['cond2']}"], dtype=bool))


def observed_matrix(ai, cci_dt_simple, mask):
    """Observed vectors (R/utils_permutation.R:13,29-63) as [R_sub, ncols] in the per-iteration column order of :580-598."""
    cond = ai["condition"]
    if not cond["is_cond"]:
        return np.asarray(cci_dt_simple["CCI_SCORE"])[mask][:, None]
    c1, c2 = cond["cond1"], cond["cond2"]
    s1 = np.asarray(cci_dt_simple[f"CCI_SCORE_{c1}"])[mask]
    s2 = np.asarray(cci_dt_simple[f"CCI_SCORE_{c2}"])[mask]
    cols = [s2 - s1, s1, s2]
    for i in range(1, ai["max_nL"] + 1):
        cols.append(np.asarray(cci_dt_simple[f"L{i}_EXPRESSION_{c2}"])[mask] - np.asarray(cci_dt_simple[f"L{i}_EXPRESSION_{c1}"])[mask])
    for j in range(1, ai["max_nR"] + 1):
        cols.append(np.asarray(cci_dt_simple[f"R{j}_EXPRESSION_{c2}"])[mask] - np.asarray(cci_dt_simple[f"R{j}_EXPRESSION_{c1}"])[mask])
    return np.stack(cols, axis=1)


def tested_row_arrays(ai, cci_dt_simple):
    """(mask, row_lri, row_emit, row_recv, observed) of the tested rows, with gene / LRI indices of the FULL inputs."""
    mask = tested_rows(ai, cci_dt_simple)
    obs = observed_matrix(ai, cci_dt_simple, mask)
    return (mask, np.asarray(cci_dt_simple["LRI_IDX"])[mask].astype(np.int32),
            np.asarray(cci_dt_simple["EMITTER_CODE"])[mask].astype(np.int32),
            np.asarray(cci_dt_simple["RECEIVER_CODE"])[mask].astype(np.int32), obs)


def build_problem(ai, cci_dt_simple):
    """Everything `.Call` receives: returns (PermutationProblem, tested-row mask)."""
    mask = tested_rows(ai, cci_dt_simple)
    obs = observed_matrix(ai, cci_dt_simple, mask)
    enc = _cci.encode_inputs(ai)
    row_lri = np.asarray(cci_dt_simple["LRI_IDX"])[mask].astype(np.int32)
    sub_gene = enc["sub_gene"]
    # gene subset of data_tr (R/utils_permutation.R:79-84): genes used by the tested rows
    used = np.unique(sub_gene[np.unique(row_lri)])
    used = used[used >= 0]
    remap = np.full(ai["data_tr"].shape[1], -1, dtype=np.int32)
    remap[used] = np.arange(len(used), dtype=np.int32)
    X = ai["data_tr"][:, used] if len(used) != ai["data_tr"].shape[1] else ai["data_tr"]
    X = X.tocsc()
    X.sort_indices()
    sub_gene = np.where(sub_gene >= 0, remap[np.maximum(sub_gene, 0)], -1).astype(np.int32)
    prob = PermutationProblem(
        n_cells=X.shape[0], n_genes=X.shape[1], n_celltypes=len(ai["cell_types"]),
        n_conditions=2 if ai["condition"]["is_cond"] else 1, max_nL=ai["max_nL"], max_nR=ai["max_nR"],
        colptr=X.indptr, rowidx=X.indices, values=X.data, ct_code=enc["ct_code"], cond_code=enc["cond_code"],
        sub_gene=sub_gene, row_lri=row_lri,
        row_emit=np.asarray(cci_dt_simple["EMITTER_CODE"])[mask].astype(np.int32),
        row_recv=np.asarray(cci_dt_simple["RECEIVER_CODE"])[mask].astype(np.int32), observed=obs)
    return prob, mask


def pvalue_column_names(ai):
    """R/utils_permutation.R:86-99, 289-300."""
    cond = ai["condition"]
    if not cond["is_cond"]:
        return ["P_VALUE"]
    # counts columns are [DE, cond1, cond2, L.., R..]
    return (["P_VALUE_DE", f"P_VALUE_{cond['cond1']}", f"P_VALUE_{cond['cond2']}"]
            + [f"L{i}_P_VALUE_DE" for i in range(1, ai["max_nL"] + 1)]
            + [f"R{j}_P_VALUE_DE" for j in range(1, ai["max_nR"] + 1)])


def run_stat_analysis(analysis_inputs, cci_dt_simple, iterations, return_distributions=False,
                      score_type="geometric_mean", verbose=False, *, seed=42, n_gpus=1, perm_cond=None, perm_ct=None,
                      batch_size=0, engine=None):
    """Mirror of run_stat_analysis(analysis_inputs, cci_dt_simple, iterations, return_distributions, score_type, verbose)
    -> {"cci_raw": table with the P_VALUE* columns added, "distributions": {...}, "stats": engine timings}.

    Keyword-only extras are out-of-band backend knobs (the reference's parameter list cannot grow):
    `seed` (the reference seeds through set.seed() in run_interaction_analysis), `n_gpus`, `batch_size`, and
    `perm_cond` / `perm_ct` = uploaded label matrices [replicates, cells] for validation against the reference's own
    shuffles. `engine`: an `Engine` already resident on a GPU with the whole `analysis_inputs` (`cci.base_problem`), e.g. the
    one that ran the observed pass: the tested rows are installed on it (`scdc_set_rows`) instead of re-uploading the matrix.
    """
    ai = analysis_inputs
    iterations = int(iterations)
    n_run = iterations_to_run(iterations, return_distributions)
    if engine is None:
        prob, mask = build_problem(ai, cci_dt_simple)
    else:
        mask, r_lri, r_emit, r_recv, obs = tested_row_arrays(ai, cci_dt_simple)
        engine.set_rows(r_lri, r_emit, r_recv, obs)
        prob = engine.problem
    if verbose:
        print(f"Performing permutation analysis ({iterations} iterations by batches of {INTERNAL_ITER}) on "
              f"{prob.n_rows} potential CCIs.")
    if engine is None:
        res = perm_counts(prob, n_run, seed=seed, score_type=score_type, n_gpus=n_gpus, batch_size=batch_size,
                          perm_cond=perm_cond, perm_ct=perm_ct, return_distributions=return_distributions)
    else:
        res = engine.run(n_run, seed=seed, score_type=score_type, batch_size=batch_size, perm_cond=perm_cond,
                         perm_ct=perm_ct, return_distributions=return_distributions)
    # p = counts / iterations (:215-259) and the merge-back (:427-469: 1 outside the tested rows, NA for absent subunits),
    # integer-coded in the shim (`scdc_merge_pvalues`, SURVEY.md "next" item f3)
    out = dict(cci_dt_simple)
    n_full = len(mask)
    raw = res.get("counts_raw")
    if raw is None:     # counts given as floats with NaN = NA (test doubles of the device call)
        c = np.asarray(res["counts"], dtype=np.float64)
        raw = np.where(np.isnan(c), np.float64(0), c).astype(np.uint64)
        raw[np.isnan(c)] = np.uint64(2**64 - 1)
    pv = merge_pvalues(n_full, np.flatnonzero(mask), raw, iterations, sub_gene=_cci.encode_inputs(ai)["sub_gene"],
                       lri_idx=np.asarray(cci_dt_simple["LRI_IDX"]), max_nL=ai["max_nL"], max_nR=ai["max_nR"])
    for k, name in enumerate(pvalue_column_names(ai)):
        out[name] = np.ascontiguousarray(pv[:, k])
    distributions = {}
    if return_distributions:                            # :304-426: rows x (iterations + 1), last column = observed
        d = res["distributions"]                        # [iterations, ncols_d, R_sub]
        obs = np.asarray(prob.observed).reshape(prob.n_rows, prob.ncols)
        cond = ai["condition"]
        names = ["DISTRIBUTIONS"] if not cond["is_cond"] else [
            "DISTRIBUTIONS_DE", f"DISTRIBUTIONS_{cond['cond1']}", f"DISTRIBUTIONS_{cond['cond2']}"]
        for k, name in enumerate(names):
            distributions[name] = np.concatenate([d[:, k, :].T, obs[:, k:k + 1]], axis=1)
    return {"cci_raw": out, "distributions": distributions, "stats": res["stats"], "tested": mask}
