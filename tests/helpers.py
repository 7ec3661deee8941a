"""Shared test helpers (fixtures decoded from the reference's .rda files live in tests/golden/)."""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_liver():
    d = np.load(os.path.join(GOLDEN, "seurat_sample_tms_liver.npz"))
    lri = dict(np.load(os.path.join(GOLDEN, "LRI_mouse_curated.npz")))
    X = sp.csc_matrix((d["x"], d["i"], d["p"]), shape=tuple(d["dim"]))   # genes x cells, log1p-normalised
    return {"X": X, "genes": d["genes"], "cells": d["cells"], "cell_type": d["cell_type"], "age_group": d["age_group"],
            "LRI_mouse": lri}


def liver_inputs(liver, mode, module, lri=None):
    """extract_analysis_inputs through `module` (oracle.oracle_np or scdiffcom_b200.preprocessing)."""
    cond = liver["age_group"] if mode == "cond" else None
    return module.extract_analysis_inputs(liver["X"], liver["genes"], liver["cells"], liver["cell_type"], cond,
                                          "YOUNG" if mode == "cond" else None, "OLD" if mode == "cond" else None,
                                          lri if lri is not None else liver["LRI_mouse"])


def liver_case(liver, mode, score_type="geometric_mean"):
    from scdiffcom_b200 import cci, permutation, preprocessing
    ai = liver_inputs(liver, mode, preprocessing)
    tmpl = cci.create_cci_template(ai)
    simple = cci.run_simple_cci_analysis(ai, tmpl, score_type=score_type)
    prob, mask = permutation.build_problem(ai, simple)
    return {"ai": ai, "simple": simple, "prob": prob, "mask": mask}


def reference_like_labels(prob, n, seed):
    """Label matrices drawn the way run_stat_iteration does (reference R/utils_permutation.R:507-510, 536-541, 552-569)
    with NumPy's Generator standing in for base::sample: uint8 (perm_cond or None, perm_ct), each [n, N]."""
    rng = np.random.default_rng(seed)
    ct = np.asarray(prob.ct_code)
    N, T = prob.n_cells, prob.n_celltypes
    perm_ct = np.empty((n, N), dtype=np.uint8)
    if prob.n_conditions == 1:
        for q in range(n):
            perm_ct[q] = rng.permutation(ct)
        return None, perm_ct
    cond = np.asarray(prob.cond_code)
    perm_cond = np.empty((n, N), dtype=np.uint8)
    first_seen = list(dict.fromkeys(ct.tolist()))      # unique(cell_type): order of first appearance
    for q in range(n):
        c = cond.copy()
        for t in first_seen:
            m = ct == t
            c[m] = rng.permutation(c[m])
        t2 = ct.copy()
        for cc in (0, 1):
            m = c == cc
            t2[m] = rng.permutation(t2[m])
        perm_cond[q], perm_ct[q] = c, t2
    return perm_cond, perm_ct


def random_problem(seed, C, N, G, T, L, nL, nR, R, empty_cols=True):
    rng = np.random.default_rng(seed)
    # labels: every (cond, ct) group non-empty
    base = np.arange(C * T).repeat(2)
    grp = np.concatenate([base, rng.integers(0, C * T, size=N - len(base))])
    grp = rng.permutation(grp)
    ct, cond = (grp % T).astype(np.int32), (grp // T).astype(np.int32)
    cols, idx, val = [0], [], []
    for g in range(G):
        d = 0.0 if (empty_cols and g % 5 == 4) else rng.uniform(0.05, 0.9)
        nz = np.flatnonzero(rng.random(N) < d)
        idx.append(nz)
        v = np.exp(rng.normal(0, 2, size=len(nz)))
        v[rng.random(len(nz)) < 0.1] = 1.0         # repeated values -> exact ties between group sums are possible
        val.append(v)
        cols.append(cols[-1] + len(nz))
    sub_gene = np.full((L, nL + nR), -1, dtype=np.int32)
    for l in range(L):
        k_l, k_r = rng.integers(1, nL + 1), rng.integers(1, nR + 1)
        genes = rng.choice(G, size=k_l + k_r, replace=(G < k_l + k_r))
        sub_gene[l, :k_l] = genes[:k_l]
        sub_gene[l, nL:nL + k_r] = genes[k_l:]
    row_lri = rng.integers(0, L, size=R).astype(np.int32)
    row_emit = rng.integers(0, T, size=R).astype(np.int32)
    row_recv = rng.integers(0, T, size=R).astype(np.int32)
    ncols = 1 if C == 1 else 3 + nL + nR
    from scdiffcom_b200 import PermutationProblem
    prob = PermutationProblem(
        n_cells=N, n_genes=G, n_celltypes=T, n_conditions=C, max_nL=nL, max_nR=nR,
        colptr=np.array(cols, dtype=np.int32), rowidx=np.concatenate(idx).astype(np.int32) if cols[-1] else np.zeros(0, np.int32),
        values=np.concatenate(val) if cols[-1] else np.zeros(0), ct_code=ct, cond_code=cond if C == 2 else None,
        sub_gene=sub_gene, row_lri=row_lri, row_emit=row_emit, row_recv=row_recv, observed=np.zeros((R, ncols)))
    return prob, rng


def with_tying_observed(prob, rng, n):
    """Observed = the values of replicate 0 (exact ties), some zeros / sign flips; NaN for absent subunits."""
    pcond, pct = reference_like_labels(prob, n, int(rng.integers(1 << 30)))
    R, C = prob.n_rows, prob.n_conditions
    if R:
        from oracle import oracle_c
        _, dist = oracle_c.perm_counts(prob, pct, pcond, return_distributions=True)
        obs = np.zeros((R, prob.ncols))
        obs[:, :dist.shape[1]] = dist[0].T
        if C == 2:
            obs[:, 3:] = rng.normal(0, 0.5, size=(R, prob.ncols - 3))
            obs[rng.random(R) < 0.2, 0] *= -1.0                      # |perm| >= |obs| is sign-blind
            absent = np.asarray(prob.sub_gene)[np.asarray(prob.row_lri)] < 0
            obs[:, 3:][absent] = np.nan
        obs[rng.random(R) < 0.1, :1] = 0.0                            # observed 0: every replicate counts
        prob.observed = obs
    return pcond, pct


