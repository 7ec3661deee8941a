"""CPU ORACLE (test infrastructure, NOT the product path) — NumPy restatement of scDiffCom's permutation test.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module; the product (`scdiffcom_b200`) never does and fails loudly without its CUDA library.

It restates, function by function, the reference's R code (all citations relative to /root/reference):

    extract_analysis_inputs   R/utils_preprocessing.R:1-50, 93-103, 220-235, 253-280, 308-324, 346-382, 402-484
    create_cci_template       R/utils_cci.R:1-30, 89-169
    aggregate_cells           R/utils_cci.R:441-462          (rowsum over cells / table(group))
    build_cci_or_drate        R/utils_cci.R:464-781          (gather, NA->0, pmin(na.rm=TRUE), score, is_detected_full)
    run_simple_cci_analysis   R/utils_cci.R:171-439
    run_stat_iteration        R/utils_permutation.R:501-603  (label shuffles, two passes, output columns)
    run_stat_analysis         R/utils_permutation.R:1-499    (row subset, batching quirk, >= counting, p-values, merge)

Third-party arithmetic that is NOT under /root/reference and is restated from its published behaviour:
  * DelayedArray::rowsum (Bioconductor, version unpinned in DESCRIPTION:59) on a dgCMatrix: for each column (gene)
    walk the stored non-zeros in ascending row (cell) order and do `out[group[row], col] += x` in double precision.
    `aggregate_cells` below reproduces that order with np.bincount (a sequential C loop over the entries).
  * base::sample / future.seed L'Ecuyer streams: NOT reproducible without R. Label vectors are therefore either
    drawn with NumPy's Generator (same distribution: uniform shuffles within the same strata) or passed in.

Parity pins (checked in tests/test_oracle_pins.py against the reference's own data and tests):
  726x468 / 73 215 nnz fixture, 1173 LRIs, 652 LRI genes, max_nL=2, max_nR=3, 29 325 template rows
  (tests/testthat/test-everything.R:281-294), 8952 / 7352 / 111 tested rows
  (docs/articles/scDiffCom-vignette.html:155-161,295-301,391-397), known-answer identities
  tests/testthat/test-everything.R:377-490 (mean, mean(x>0), sqrt(mean*mean) from raw data, tol 1.5e-8).
Permutation p-values themselves are unpinned in the reference (no golden p-values exist).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

NA = ""  # NA_character_ in the decoded fixtures


# --------------------------------------------------------------------------------------------------------------
# R/utils_preprocessing.R
# --------------------------------------------------------------------------------------------------------------
def extract_analysis_inputs(expr_gc, genes, cell_ids, cell_type, condition, cond1_name, cond2_name,
                            LRI_table, slot="data", log_scale=False, threshold_min_cells=5):
    """expr_gc: genes x cells scipy CSC (the Seurat slot). Returns the `analysis_inputs` list as a dict
    (R/utils_preprocessing.R:41-49); `data_tr` is cells x genes CSC restricted to LRI genes."""
    expr_gc = sp.csc_matrix(expr_gc, dtype=np.float64, copy=True)
    if slot == "data" and not log_scale:          # :93-103
        expr_gc.data = np.expm1(expr_gc.data)
    if slot == "counts" and log_scale:            # :104-111
        expr_gc.data = np.log1p(expr_gc.data)
    cell_type = np.array([c.replace("_", "-") for c in cell_type])   # :220-229
    is_cond = condition is not None
    # filter_celltypes :253-280 (table() sorts level names)
    levels = sorted(set(cell_type.tolist()))
    keep = []
    for ct in levels:
        m = cell_type == ct
        if is_cond:
            conds = sorted(set(np.asarray(condition).tolist()))
            ok = all(int(np.sum(m & (np.asarray(condition) == c))) >= threshold_min_cells for c in conds)
        else:
            ok = int(m.sum()) >= threshold_min_cells
        if ok:
            keep.append(ct)
    if len(keep) < 2:
        raise ValueError("Inputs must contain at least 2 cell-types")
    cell_mask = np.isin(cell_type, keep)          # :234-235
    expr_gc = expr_gc[:, np.flatnonzero(cell_mask)]
    metadata = {"cell_id": np.asarray(cell_ids)[cell_mask], "cell_type": cell_type[cell_mask]}
    if is_cond:
        metadata["condition"] = np.asarray(condition)[cell_mask]
    # extract_LRI_inputs :282-400
    cols = ["LRI", "LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3"]
    rows = list(dict.fromkeys(zip(*[LRI_table[c].tolist() for c in cols])))   # unique(), first occurrence order
    geneset = set(np.asarray(genes).tolist())
    ok_or_na = lambda g: g == NA or g in geneset
    rows = [r for r in rows if r[1] in geneset and r[3] in geneset and ok_or_na(r[2]) and ok_or_na(r[4]) and ok_or_na(r[5])]
    LRI = {c: np.array([r[k] for r in rows], dtype=np.str_) for k, c in enumerate(cols)}
    lri_genes = set()
    for c in cols[1:]:
        lri_genes.update(g for g in LRI[c].tolist() if g != NA)
    gene_mask = np.array([g in lri_genes for g in np.asarray(genes).tolist()])
    data_keep = expr_gc[np.flatnonzero(gene_mask), :]
    max_nL = 1 if np.all(LRI["LIGAND_2"] == NA) else 2                      # :354-359
    if np.all(LRI["RECEPTOR_2"] == NA) and np.all(LRI["RECEPTOR_3"] == NA):  # :360-381
        max_nR = 1
    elif np.all(LRI["RECEPTOR_2"] == NA):
        raise ValueError("RECEPTOR_3 without RECEPTOR_2")
    elif np.all(LRI["RECEPTOR_3"] == NA):
        max_nR = 2
    else:
        max_nR = 3
    if max_nL == 1:
        LRI.pop("LIGAND_2")
    if max_nR < 3:
        LRI.pop("RECEPTOR_3")
    if max_nR < 2:
        LRI.pop("RECEPTOR_2")
    cond = {"is_cond": is_cond, "is_samp": False}
    if is_cond:                                   # :435-453
        cond["cond1"], cond["cond2"] = cond1_name, cond2_name
        assert {cond1_name, cond2_name} == set(metadata["condition"].tolist())
    return {
        "data_tr": sp.csc_matrix(data_keep.T),    # cells x genes (DelayedArray::t), :42
        "genes": np.asarray(genes)[gene_mask],
        "metadata": metadata,
        "cell_types": keep,
        "LRI": LRI,
        "max_nL": max_nL,
        "max_nR": max_nR,
        "condition": cond,
    }


# --------------------------------------------------------------------------------------------------------------
# R/utils_cci.R
# --------------------------------------------------------------------------------------------------------------
def create_cci_template(ai):
    """CJ(EMITTER, RECEIVER, LRI) joined with the LRI table and NCELLS (R/utils_cci.R:1-30, 89-169)."""
    cts = sorted(ai["cell_types"])
    lris = sorted(ai["LRI"]["LRI"].tolist())
    lri_pos = {name: k for k, name in enumerate(ai["LRI"]["LRI"].tolist())}
    T, L = len(cts), len(lris)
    tmpl = {
        "EMITTER_CELLTYPE": np.repeat(np.array(cts), T * L),
        "RECEIVER_CELLTYPE": np.tile(np.repeat(np.array(cts), L), T),
        "LRI": np.tile(np.array(lris), T * T),
    }
    idx = np.array([lri_pos[x] for x in tmpl["LRI"].tolist()])
    for c in ai["LRI"]:
        if c != "LRI":
            tmpl[c] = ai["LRI"][c][idx]
    md = ai["metadata"]
    if not ai["condition"]["is_cond"]:
        n = {ct: int(np.sum(md["cell_type"] == ct)) for ct in cts}
        tmpl["NCELLS_EMITTER"] = np.array([n[c] for c in tmpl["EMITTER_CELLTYPE"].tolist()])
        tmpl["NCELLS_RECEIVER"] = np.array([n[c] for c in tmpl["RECEIVER_CELLTYPE"].tolist()])
    else:
        for cname in (ai["condition"]["cond1"], ai["condition"]["cond2"]):
            n = {ct: int(np.sum((md["cell_type"] == ct) & (md["condition"] == cname))) for ct in cts}
            tmpl[f"NCELLS_EMITTER_{cname}"] = np.array([n[c] for c in tmpl["EMITTER_CELLTYPE"].tolist()])
            tmpl[f"NCELLS_RECEIVER_{cname}"] = np.array([n[c] for c in tmpl["RECEIVER_CELLTYPE"].tolist()])
    return tmpl


def aggregate_cells(data_tr, metadata, is_cond):
    """rowsum(data_tr, group, reorder=TRUE) / table(group) (R/utils_cci.R:441-462).
    Returns (group names sorted, K x G matrix). Sums run over the stored entries of each gene column in ascending
    cell order, in double precision (np.bincount is a sequential loop over its input)."""
    if not is_cond:
        group = np.asarray(metadata["cell_type"])
    else:
        group = np.char.add(np.char.add(np.asarray(metadata["condition"]), "_"), np.asarray(metadata["cell_type"]))
    names, code = np.unique(group, return_inverse=True)   # sorted, like reorder=TRUE / table()
    K = len(names)
    X = sp.csc_matrix(data_tr)
    X.sort_indices()
    G = X.shape[1]
    gene_of_entry = np.repeat(np.arange(G, dtype=np.int64), np.diff(X.indptr))
    flat = code[X.indices].astype(np.int64) * G + gene_of_entry
    sums = np.bincount(flat, weights=X.data, minlength=K * G).reshape(K, G)
    counts = np.bincount(code, minlength=K).astype(np.float64)
    return names, sums / counts[:, None]


def _pmin_na_rm(cols):
    """pmin(..., na.rm=TRUE) over a list of float arrays with NaN = NA."""
    out = cols[0].copy()
    for c in cols[1:]:
        out = np.where(np.isnan(out), c, np.where(np.isnan(c), out, np.minimum(out, c)))
    return out


def build_cci_or_drate(names, averaged, genes, tmpl, max_nL, max_nR, cond, cci_or_drate, score_type,
                       threshold_min_cells=None, threshold_pct=None):
    """R/utils_cci.R:464-781. `averaged` is K x G with row names `names`."""
    tag = "EXPRESSION" if cci_or_drate == "cci" else "DETECTION_RATE"
    gene_pos = {g: k for k, g in enumerate(np.asarray(genes).tolist())}
    row_of = {n: k for k, n in enumerate(names.tolist())}
    suffixes = [""] if not cond["is_cond"] else [f"_{cond['cond1']}", f"_{cond['cond2']}"]
    prefixes = [""] if not cond["is_cond"] else [f"{cond['cond1']}_", f"{cond['cond2']}_"]
    out = dict(tmpl)
    n_rows = len(tmpl["LRI"])

    def gather(gene_col, ct_col, prefix):
        res = np.full(n_rows, np.nan)
        g = np.array([gene_pos.get(x, -1) for x in tmpl[gene_col].tolist()])
        k = np.array([row_of.get(prefix + c, -1) for c in tmpl[ct_col].tolist()])
        present = np.array([x != NA for x in tmpl[gene_col].tolist()])
        ok = present & (g >= 0) & (k >= 0)
        res[ok] = averaged[k[ok], g[ok]]
        # dt[is.na(dt)] <- 0 (:555) applies to the melted means; a present gene with a missing group mean is 0
        res[present & ~ok] = 0.0
        return res

    for sfx, pfx in zip(suffixes, prefixes):
        for i in range(1, max_nL + 1):
            out[f"L{i}_{tag}{sfx}"] = gather(f"LIGAND_{i}", "EMITTER_CELLTYPE", pfx)
        for j in range(1, max_nR + 1):
            out[f"R{j}_{tag}{sfx}"] = gather(f"RECEPTOR_{j}", "RECEIVER_CELLTYPE", pfx)
    for sfx in suffixes:
        minL = _pmin_na_rm([out[f"L{i}_{tag}{sfx}"] for i in range(1, max_nL + 1)])
        minR = _pmin_na_rm([out[f"R{j}_{tag}{sfx}"] for j in range(1, max_nR + 1)])
        if cci_or_drate == "cci":
            if score_type == "geometric_mean":
                out[f"CCI_SCORE{sfx}"] = np.sqrt(minL * minR)          # :605-658
            elif score_type == "arithmetic_mean":
                out[f"CCI_SCORE{sfx}"] = (minL + minR) / 2             # :659-712
            else:
                raise ValueError(score_type)
        else:                                                          # is_detected_full :783-802
            ne = tmpl[f"NCELLS_EMITTER{sfx}"]
            nr = tmpl[f"NCELLS_RECEIVER{sfx}"]
            out[f"IS_CCI_EXPRESSED{sfx}"] = (
                (minL >= threshold_pct) & (minL * ne >= threshold_min_cells)
                & (minR >= threshold_pct) & (minR * nr >= threshold_min_cells)
            )
    return out


def run_simple_cci_analysis(ai, tmpl, score_type="geometric_mean", threshold_min_cells=5, threshold_pct=0.1,
                            compute_fast=False, log_scale=False):
    """R/utils_cci.R:171-439."""
    cond = ai["condition"]
    names, avg = aggregate_cells(ai["data_tr"], ai["metadata"], cond["is_cond"])
    cci = build_cci_or_drate(names, avg, ai["genes"], tmpl, ai["max_nL"], ai["max_nR"], cond, "cci", score_type)
    if compute_fast:
        if not cond["is_cond"]:
            return cci["CCI_SCORE"]
        c1, c2 = cond["cond1"], cond["cond2"]
        res = {"cond1": cci[f"CCI_SCORE_{c1}"], "cond2": cci[f"CCI_SCORE_{c2}"]}
        for i in range(1, ai["max_nL"] + 1):
            res[f"L{i}_DIFF_EXPR"] = cci[f"L{i}_EXPRESSION_{c2}"] - cci[f"L{i}_EXPRESSION_{c1}"]
        for j in range(1, ai["max_nR"] + 1):
            res[f"R{j}_DIFF_EXPR"] = cci[f"R{j}_EXPRESSION_{c2}"] - cci[f"R{j}_EXPRESSION_{c1}"]
        return res
    X = sp.csc_matrix(ai["data_tr"], copy=True)
    X.data = 1.0 * (X.data > 0)                                        # :262-266
    names_d, dr = aggregate_cells(X, ai["metadata"], cond["is_cond"])
    out = build_cci_or_drate(names_d, dr, ai["genes"], cci, ai["max_nL"], ai["max_nR"], cond, "drate", score_type,
                             threshold_min_cells=threshold_min_cells, threshold_pct=threshold_pct)
    if cond["is_cond"]:                                                # :284-437
        c1, c2 = cond["cond1"], cond["cond2"]
        sub = [f"L{i}" for i in range(1, ai["max_nL"] + 1)] + [f"R{j}" for j in range(1, ai["max_nR"] + 1)]
        gene_cols = [f"LIGAND_{i}" for i in range(1, ai["max_nL"] + 1)] + [f"RECEPTOR_{j}" for j in range(1, ai["max_nR"] + 1)]
        with np.errstate(divide="ignore", invalid="ignore"):
            if log_scale:                                              # :299-352: plain differences, nothing replaced
                lfc = out[f"CCI_SCORE_{c2}"] - out[f"CCI_SCORE_{c1}"]
                for name in sub:
                    out[f"{name}_LOGFC"] = out[f"{name}_EXPRESSION_{c2}"] - out[f"{name}_EXPRESSION_{c1}"]
            else:                                                      # :353-433: log of the ratio, NaN (0/0) -> 0;
                lfc = np.log(out[f"CCI_SCORE_{c2}"] / out[f"CCI_SCORE_{c1}"])
                lfc = np.where(np.isnan(lfc), 0.0, lfc)
                for name, gcol in zip(sub, gene_cols):
                    t = np.log(out[f"{name}_EXPRESSION_{c2}"] / out[f"{name}_EXPRESSION_{c1}"])
                    absent = np.array([x == NA for x in tmpl[gcol].tolist()])   # NA subunit: is.nan(NA) is FALSE, NA is kept
                    out[f"{name}_LOGFC"] = np.where(absent, np.nan, np.where(np.isnan(t), 0.0, t))
        out["LOGFC"] = lfc
        out["LOGFC_ABS"] = np.abs(lfc)
    return out


# --------------------------------------------------------------------------------------------------------------
# R/utils_permutation.R
# --------------------------------------------------------------------------------------------------------------
def shuffle_labels(ai, rng):
    """The label shuffles of run_stat_iteration (R/utils_permutation.R:507-510, 536-541, 552-569) with a NumPy
    Generator standing in for base::sample. Returns the permuted metadata dict(s):
    detection: (meta_ct,) ; differential: (meta_cond, meta_ct)."""
    md = ai["metadata"]
    cond = ai["condition"]
    if not cond["is_cond"]:
        meta_ct = dict(md)
        meta_ct["cell_type"] = rng.permutation(md["cell_type"])        # sample(meta_ct$cell_type) :509
        return (meta_ct,)
    meta_cond = {k: v.copy() for k, v in md.items()}
    # unique(cell_type) in order of first appearance :537
    for x in list(dict.fromkeys(md["cell_type"].tolist())):
        m = meta_cond["cell_type"] == x
        meta_cond["condition"][m] = rng.permutation(meta_cond["condition"][m])     # :538-539
    meta_ct = {k: v.copy() for k, v in meta_cond.items()}              # starts from the permuted conditions :552
    for cname in (cond["cond1"], cond["cond2"]):                       # :553-568
        m = meta_ct["condition"] == cname
        meta_ct["cell_type"][m] = rng.permutation(meta_ct["cell_type"][m])
    return meta_cond, meta_ct


def run_stat_iteration(ai, tmpl, score_type, rng=None, metas=None):
    """R/utils_permutation.R:501-603. Returns a vector (detection) or an R_sub x (3+nL+nR) matrix."""
    if metas is None:
        metas = shuffle_labels(ai, rng)
    cond = ai["condition"]
    if not cond["is_cond"]:
        a = dict(ai); a["metadata"] = metas[0]
        return run_simple_cci_analysis(a, tmpl, score_type, compute_fast=True)
    a = dict(ai); a["metadata"] = metas[0]
    permcond = run_simple_cci_analysis(a, tmpl, score_type, compute_fast=True)       # :543-551
    a = dict(ai); a["metadata"] = metas[1]
    permct = run_simple_cci_analysis(a, tmpl, score_type, compute_fast=True)         # :570-578
    cols = [permcond["cond2"] - permcond["cond1"], permct["cond1"], permct["cond2"]]  # :579-585
    cols += [permcond[f"L{i}_DIFF_EXPR"] for i in range(1, ai["max_nL"] + 1)]
    cols += [permcond[f"R{j}_DIFF_EXPR"] for j in range(1, ai["max_nR"] + 1)]
    return np.stack(cols, axis=1)


def tested_rows_and_observed(ai, cci_dt_simple):
    """Row subset and observed vectors of run_stat_analysis (R/utils_permutation.R:11-64).
    Returns (row mask over the template, observed matrix R_sub x ncols in the order of :580-598)."""
    cond = ai["condition"]
    if not cond["is_cond"]:
        mask = cci_dt_simple["IS_CCI_EXPRESSED"]
        return mask, cci_dt_simple["CCI_SCORE"][mask][:, None]
    c1, c2 = cond["cond1"], cond["cond2"]
    mask = cci_dt_simple[f"IS_CCI_EXPRESSED_{c1}"] | cci_dt_simple[f"IS_CCI_EXPRESSED_{c2}"]
    s1 = cci_dt_simple[f"CCI_SCORE_{c1}"][mask]
    s2 = cci_dt_simple[f"CCI_SCORE_{c2}"][mask]
    cols = [s2 - s1, s1, s2]
    for i in range(1, ai["max_nL"] + 1):
        cols.append(cci_dt_simple[f"L{i}_EXPRESSION_{c2}"][mask] - cci_dt_simple[f"L{i}_EXPRESSION_{c1}"][mask])
    for j in range(1, ai["max_nR"] + 1):
        cols.append(cci_dt_simple[f"R{j}_EXPRESSION_{c2}"][mask] - cci_dt_simple[f"R{j}_EXPRESSION_{c1}"][mask])
    return mask, np.stack(cols, axis=1)


def restrict_to_tested(ai, tmpl, mask):
    """sub_cci_template_iter and the gene subset of data_tr (R/utils_permutation.R:75-85,114)."""
    lr_cols = [f"LIGAND_{i}" for i in range(1, ai["max_nL"] + 1)] + [f"RECEPTOR_{j}" for j in range(1, ai["max_nR"] + 1)]
    keep_cols = ["LRI", "EMITTER_CELLTYPE", "RECEIVER_CELLTYPE"] + lr_cols
    sub = {c: tmpl[c][mask] for c in keep_cols}
    used = set()
    for c in lr_cols:
        used.update(sub[c].tolist())
    gmask = np.array([g in used for g in ai["genes"].tolist()])
    a = dict(ai)
    a["data_tr"] = sp.csc_matrix(ai["data_tr"])[:, np.flatnonzero(gmask)]
    a["genes"] = ai["genes"][gmask]
    return a, sub


def exceedance_counts(perm, obs, is_cond):
    """One replicate's contribution to the counters (R/utils_permutation.R:139-211): `>=`, two-sided abs() for the DE
    and per-subunit columns, one-sided for detection/cond1/cond2. NaN (NA subunit) compares FALSE."""
    with np.errstate(invalid="ignore"):
        if not is_cond:
            return (perm[:, None] >= obs).astype(np.int64)
        out = np.abs(perm) >= np.abs(obs)
        out[:, 1] = perm[:, 1] >= obs[:, 1]
        out[:, 2] = perm[:, 2] >= obs[:, 2]
    return out.astype(np.int64)


def iterations_actually_run(iterations, internal_iter=1000):
    """Batching quirk of R/utils_permutation.R:100-128: floor(iterations/1000) batches, the last one running
    iterations %% 1000 (or 1000) replicates. Equals `iterations` when iterations <= 1000 or a multiple of 1000."""
    if iterations <= internal_iter:
        return iterations
    n_broad = iterations // internal_iter
    last = iterations % internal_iter or internal_iter
    return (n_broad - 1) * internal_iter + last


def run_stat_analysis(ai, cci_dt_simple, iterations, score_type="geometric_mean", rng=None, label_metas=None,
                      return_distributions=False):
    """R/utils_permutation.R:1-499 (counting branch and distributions branch). `label_metas`: optional list of
    permuted metadata tuples, one per replicate (uploaded-permutation mode). Returns dict with `mask`, `counts`
    (R_sub x ncols int64), `pvals`, `observed`, and optionally `distributions` (R_sub x ncols x iterations)."""
    is_cond = ai["condition"]["is_cond"]
    mask, obs = tested_rows_and_observed(ai, cci_dt_simple)
    a, sub = restrict_to_tested(ai, cci_dt_simple, mask)
    n_run = iterations if return_distributions else iterations_actually_run(iterations)
    counts = np.zeros(obs.shape, dtype=np.int64)
    dist = np.empty(obs.shape + (n_run,)) if return_distributions else None
    for it in range(n_run):
        metas = label_metas[it] if label_metas is not None else None
        perm = run_stat_iteration(a, sub, score_type, rng=rng, metas=metas)
        counts += exceedance_counts(perm, obs, is_cond)
        if dist is not None:
            dist[:, :, it] = perm[:, None] if not is_cond else perm
    pvals = counts / iterations                                         # :215-259 (divides by `iterations`)
    # absent subunits: NA (:442-469)
    pvals = pvals.astype(np.float64)
    pvals[np.isnan(obs)] = np.nan
    return {"mask": mask, "observed": obs, "counts": counts, "pvals": pvals, "distributions": dist,
            "inputs": a, "template": sub}
