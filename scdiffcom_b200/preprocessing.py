"""Host mirror of the reference's input extraction (R/utils_preprocessing.R) for callers without R.

In the R integration these steps stay in R (they run once per analysis; SURVEY.md section 2, component 4); this module
exists so that the Python tests and benchmarks can build `analysis_inputs` exactly the way the reference does.
Strings are turned into integer codes here, once; NA_character_ is the empty string.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

LRI_COLUMNS = ("LRI", "LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3")
NA = ""


def filter_celltypes(cell_type, condition, threshold_min_cells):
    """R/utils_preprocessing.R:253-280: cell types with >= threshold cells (in every condition), sorted by name."""
    names, code = np.unique(cell_type, return_inverse=True)
    if condition is None:
        ok = np.bincount(code, minlength=len(names)) >= threshold_min_cells
    else:
        cnames, ccode = np.unique(condition, return_inverse=True)
        tab = np.bincount(code * len(cnames) + ccode, minlength=len(names) * len(cnames)).reshape(len(names), -1)
        ok = (tab >= threshold_min_cells).all(axis=1)
    kept = names[ok].tolist()
    if len(kept) < 2:
        raise ValueError(f"Inputs must contain at least 2 cell-types with at least {threshold_min_cells} cells (in each condition).")
    return kept


def extract_analysis_inputs(expr_genes_by_cells, genes, cell_ids, cell_type, condition=None, cond1_name=None,
                            cond2_name=None, LRI_table=None, slot="data", log_scale=False, threshold_min_cells=5):
    """Mirror of extract_analysis_inputs (R/utils_preprocessing.R:1-50) for a genes x cells sparse matrix plus metadata
    vectors instead of a Seurat object. Returns the `analysis_inputs` list as a dict; `data_tr` is cells x genes CSC."""
    X = sp.csc_matrix(expr_genes_by_cells, dtype=np.float64, copy=True)
    genes = np.asarray(genes)
    if slot == "data" and not log_scale:      # :93-103
        X.data = np.expm1(X.data)
    elif slot == "counts" and log_scale:      # :104-111
        X.data = np.log1p(X.data)
    cell_type = np.char.replace(np.asarray(cell_type, dtype=np.str_), "_", "-")   # :220-229
    is_cond = condition is not None
    if is_cond:
        condition = np.asarray(condition, dtype=np.str_)
        present = set(np.unique(condition).tolist())
        if len(present) != 2:
            raise ValueError(f"condition must contain exactly two groups ({len(present)} supplied)")
        if {cond1_name, cond2_name} != present:
            raise ValueError("Either 'cond1_name' or 'cond2_name' does not match with the content of the condition column")
    cell_types = filter_celltypes(cell_type, condition, threshold_min_cells)
    keep = np.isin(cell_type, cell_types)
    X = X[:, np.flatnonzero(keep)]
    metadata = {"cell_id": np.asarray(cell_ids)[keep], "cell_type": cell_type[keep]}
    if is_cond:
        metadata["condition"] = condition[keep]
    # extract_LRI_inputs :282-400
    table = np.stack([np.asarray(LRI_table[c], dtype=np.str_) for c in LRI_COLUMNS], axis=1)
    _, first = np.unique(table, axis=0, return_index=True)
    table = table[np.sort(first)]              # unique(), keeping first occurrences in order
    in_data = lambda col: np.isin(col, genes)
    ok = in_data(table[:, 1]) & in_data(table[:, 3])
    for k in (2, 4, 5):
        ok &= in_data(table[:, k]) | (table[:, k] == NA)
    table = table[ok]
    if len(np.unique(table[:, 0])) == 0:
        raise ValueError("There are no genes from 'LRI_table' in the data.")
    lri_genes = np.unique(table[:, 1:])
    lri_genes = lri_genes[lri_genes != NA]
    gmask = np.isin(genes, lri_genes)
    X = X[np.flatnonzero(gmask), :]
    LRI = {c: table[:, k] for k, c in enumerate(LRI_COLUMNS)}
    max_nL = 1 if np.all(LRI["LIGAND_2"] == NA) else 2
    if np.all(LRI["RECEPTOR_2"] == NA) and np.all(LRI["RECEPTOR_3"] == NA):
        max_nR = 1
    elif np.all(LRI["RECEPTOR_2"] == NA):
        raise ValueError("'LRI_table' must not contain only NA in column 'RECEPTOR_2' and non-NA in column 'RECEPTOR_3'.")
    elif np.all(LRI["RECEPTOR_3"] == NA):
        max_nR = 2
    else:
        max_nR = 3
    for c in ["LIGAND_2"] * (max_nL < 2) + ["RECEPTOR_3"] * (max_nR < 3) + ["RECEPTOR_2"] * (max_nR < 2):
        LRI.pop(c)
    cond = {"is_cond": is_cond, "is_samp": False}
    if is_cond:
        cond.update(cond1=cond1_name, cond2=cond2_name)
    data_tr = sp.csc_matrix(X.T)
    data_tr.sort_indices()
    return {"data_tr": data_tr, "genes": genes[gmask], "metadata": metadata, "cell_types": cell_types, "LRI": LRI,
            "max_nL": max_nL, "max_nR": max_nR, "condition": cond}
