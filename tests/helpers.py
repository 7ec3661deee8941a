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
