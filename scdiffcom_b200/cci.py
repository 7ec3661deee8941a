"""Host mirror of the observed pass (reference R/utils_cci.R): CCI template, group means, detection rates, scores.

Stays in R in the real integration (runs once; SURVEY.md section 8f "next" item f2 moves it to the GPU). Everything is
integer-coded: a template row is (EMITTER_CODE, RECEIVER_CODE, LRI_IDX) and tables are dicts of NumPy columns named as in
the reference (`CCI_SCORE_<cond>`, `L1_EXPRESSION_<cond>`, `IS_CCI_EXPRESSED_<cond>`, ...).
"""
from __future__ import annotations

import numpy as np

NA = ""


def subunit_columns(ai):
    return [f"LIGAND_{i}" for i in range(1, ai["max_nL"] + 1)] + [f"RECEPTOR_{j}" for j in range(1, ai["max_nR"] + 1)]


def encode_inputs(ai):
    """Integer codes used on both sides of the boundary (SURVEY.md section 8b "Marshaling done in R before the call")."""
    cts = list(ai["cell_types"])
    ct_pos = {c: k for k, c in enumerate(cts)}
    ct_code = np.fromiter((ct_pos[c] for c in ai["metadata"]["cell_type"].tolist()), dtype=np.int32)
    cond = ai["condition"]
    cond_code = None
    if cond["is_cond"]:
        cond_code = (np.asarray(ai["metadata"]["condition"]) == cond["cond2"]).astype(np.int32)
    gene_pos = {g: k for k, g in enumerate(np.asarray(ai["genes"]).tolist())}
    cols = subunit_columns(ai)
    sub_gene = np.full((len(ai["LRI"]["LRI"]), len(cols)), -1, dtype=np.int32)
    for s, c in enumerate(cols):
        sub_gene[:, s] = [gene_pos.get(g, -1) if g != NA else -1 for g in ai["LRI"][c].tolist()]
    return {"ct_code": ct_code, "cond_code": cond_code, "sub_gene": sub_gene}


def create_cci_template(ai):
    """CJ(EMITTER_CELLTYPE, RECEIVER_CELLTYPE, LRI) (R/utils_cci.R:1-30): rows ordered by emitter, receiver, LRI name."""
    T = len(ai["cell_types"])
    order = np.argsort(np.asarray(ai["LRI"]["LRI"]), kind="stable").astype(np.int32)
    L = len(order)
    enc = encode_inputs(ai)
    tmpl = {
        "EMITTER_CODE": np.repeat(np.arange(T, dtype=np.int32), T * L),
        "RECEIVER_CODE": np.tile(np.repeat(np.arange(T, dtype=np.int32), L), T),
        "LRI_IDX": np.tile(order, T * T),
    }
    K = T * (2 if ai["condition"]["is_cond"] else 1)
    gcode = enc["ct_code"] + (T * enc["cond_code"] if enc["cond_code"] is not None else 0)
    ncells = np.bincount(gcode, minlength=K)      # add_cell_number, R/utils_cci.R:89-169
    cond = ai["condition"]
    if not cond["is_cond"]:
        tmpl["NCELLS_EMITTER"] = ncells[tmpl["EMITTER_CODE"]]
        tmpl["NCELLS_RECEIVER"] = ncells[tmpl["RECEIVER_CODE"]]
    else:
        for c, name in enumerate((cond["cond1"], cond["cond2"])):
            tmpl[f"NCELLS_EMITTER_{name}"] = ncells[c * T + tmpl["EMITTER_CODE"]]
            tmpl[f"NCELLS_RECEIVER_{name}"] = ncells[c * T + tmpl["RECEIVER_CODE"]]
    return tmpl


def aggregate_cells(data_tr, group_code, K, detection=False):
    """rowsum(data_tr, group) / table(group) (R/utils_cci.R:441-462) -> [K, G]; ascending-cell fp64 sums per (gene, group)."""
    G = data_tr.shape[1]
    gene_of_entry = np.repeat(np.arange(G, dtype=np.int64), np.diff(data_tr.indptr))
    w = (data_tr.data > 0).astype(np.float64) if detection else data_tr.data
    sums = np.bincount(group_code[data_tr.indices].astype(np.int64) * G + gene_of_entry, weights=w, minlength=K * G)
    return sums.reshape(K, G) / np.bincount(group_code, minlength=K).astype(np.float64)[:, None]


def _pmin(cols):
    out = cols[0]
    for c in cols[1:]:
        out = np.fmin(out, c)   # fmin ignores NaN = pmin(na.rm = TRUE)
    return out


def base_problem(ai):
    """`PermutationProblem` holding the whole `analysis_inputs` (all LRI genes, no tested rows yet): what a resident
    engine needs for the observed pass; the tested rows are installed later with `Engine.set_rows`."""
    from .engine import PermutationProblem
    enc = encode_inputs(ai)
    X = ai["data_tr"].tocsc()
    X.sort_indices()
    C = 2 if ai["condition"]["is_cond"] else 1
    ncols = 1 if C == 1 else 3 + ai["max_nL"] + ai["max_nR"]
    return PermutationProblem(
        n_cells=X.shape[0], n_genes=X.shape[1], n_celltypes=len(ai["cell_types"]), n_conditions=C,
        max_nL=ai["max_nL"], max_nR=ai["max_nR"], colptr=X.indptr, rowidx=X.indices, values=X.data,
        ct_code=enc["ct_code"], cond_code=enc["cond_code"], sub_gene=enc["sub_gene"],
        row_lri=np.zeros(0, np.int32), row_emit=np.zeros(0, np.int32), row_recv=np.zeros(0, np.int32),
        observed=np.zeros((0, ncols)))


def run_simple_cci_analysis(ai, tmpl, score_type="geometric_mean", threshold_min_cells=5, threshold_pct=0.1,
                            log_scale=False, engine=None):
    """Observed pass (`compute_fast = FALSE`, R/utils_cci.R:171-439): expression, detection rate, CCI_SCORE*,
    IS_CCI_EXPRESSED*, LOGFC. Returns a new table (dict of columns).

    engine = None: NumPy on the host (what R does today). engine = an `Engine` built from `base_problem(ai)`: group means,
    detection rates, gathers, scores and IS_CCI_EXPRESSED come from the GPU (`scdc_observed_pass`, SURVEY.md "next" f2);
    only the LOGFC (a log) is finished on the host."""
    if engine is not None:
        return _run_simple_cci_analysis_b200(ai, tmpl, engine, score_type, threshold_min_cells, threshold_pct, log_scale)
    enc = encode_inputs(ai)
    T = len(ai["cell_types"])
    cond = ai["condition"]
    C = 2 if cond["is_cond"] else 1
    gcode = enc["ct_code"] + (T * enc["cond_code"] if C == 2 else 0)
    X = ai["data_tr"]
    means = aggregate_cells(X, gcode, C * T)
    drate = aggregate_cells(X, gcode, C * T, detection=True)
    out = dict(tmpl)
    sg = enc["sub_gene"][tmpl["LRI_IDX"]]          # [rows, nL+nR]
    nL, nR = ai["max_nL"], ai["max_nR"]
    sfx = [""] if C == 1 else [f"_{cond['cond1']}", f"_{cond['cond2']}"]
    for c in range(C):
        for tag, tab in (("EXPRESSION", means), ("DETECTION_RATE", drate)):
            cols = []
            for s in range(nL + nR):
                grp = c * T + (tmpl["EMITTER_CODE"] if s < nL else tmpl["RECEIVER_CODE"])
                v = np.where(sg[:, s] >= 0, tab[grp, np.maximum(sg[:, s], 0)], np.nan)
                name = (f"L{s + 1}" if s < nL else f"R{s - nL + 1}") + f"_{tag}{sfx[c]}"
                out[name] = v
                cols.append(v)
            minL, minR = _pmin(cols[:nL]), _pmin(cols[nL:])
            if tag == "EXPRESSION":
                if score_type == "geometric_mean":
                    out[f"CCI_SCORE{sfx[c]}"] = np.sqrt(minL * minR)
                elif score_type == "arithmetic_mean":
                    out[f"CCI_SCORE{sfx[c]}"] = (minL + minR) / 2
                else:
                    raise ValueError(f"unknown score_type {score_type!r}")
            else:   # is_detected_full, R/utils_cci.R:783-802
                ne, nr = tmpl[f"NCELLS_EMITTER{sfx[c]}"], tmpl[f"NCELLS_RECEIVER{sfx[c]}"]
                out[f"IS_CCI_EXPRESSED{sfx[c]}"] = ((minL >= threshold_pct) & (minL * ne >= threshold_min_cells)
                                                    & (minR >= threshold_pct) & (minR * nr >= threshold_min_cells))
    if C == 2:
        s1, s2 = out[f"CCI_SCORE{sfx[0]}"], out[f"CCI_SCORE{sfx[1]}"]
        with np.errstate(divide="ignore", invalid="ignore"):
            lfc = (s2 - s1) if log_scale else np.log(s2 / s1)
        out["LOGFC"] = np.where(np.isnan(lfc), 0.0, lfc)
        out["LOGFC_ABS"] = np.abs(out["LOGFC"])
    return out


def _run_simple_cci_analysis_b200(ai, tmpl, engine, score_type, threshold_min_cells, threshold_pct, log_scale):
    cond = ai["condition"]
    C = 2 if cond["is_cond"] else 1
    nL, nR = ai["max_nL"], ai["max_nR"]
    res = engine.observed_pass(tmpl["LRI_IDX"], tmpl["EMITTER_CODE"], tmpl["RECEIVER_CODE"], score_type=score_type,
                               threshold_min_cells=threshold_min_cells, threshold_pct=threshold_pct)
    out = dict(tmpl)
    sfx = [""] if C == 1 else [f"_{cond['cond1']}", f"_{cond['cond2']}"]
    for c in range(C):
        for s in range(nL + nR):
            name = f"L{s + 1}" if s < nL else f"R{s - nL + 1}"
            out[f"{name}_EXPRESSION{sfx[c]}"] = np.ascontiguousarray(res["expression"][:, c, s])
            out[f"{name}_DETECTION_RATE{sfx[c]}"] = np.ascontiguousarray(res["detection_rate"][:, c, s])
        out[f"CCI_SCORE{sfx[c]}"] = np.ascontiguousarray(res["score"][:, c])
        out[f"IS_CCI_EXPRESSED{sfx[c]}"] = np.ascontiguousarray(res["is_expressed"][:, c])
    if C == 2:
        s1, s2 = out[f"CCI_SCORE{sfx[0]}"], out[f"CCI_SCORE{sfx[1]}"]
        with np.errstate(divide="ignore", invalid="ignore"):
            lfc = (s2 - s1) if log_scale else np.log(s2 / s1)
        out["LOGFC"] = np.where(np.isnan(lfc), 0.0, lfc)
        out["LOGFC_ABS"] = np.abs(out["LOGFC"])
    return out
