"""Decode the reference's bundled `.rda` data into `.npz` golden fixtures (run ONCE, in the build container).

    python tests/golden/make_fixtures.py            # reads /root/reference/data, writes tests/golden/*.npz

`/root/reference` does not exist on the GPU box, so the decoded arrays are committed next to this script.
Nothing is computed here: the arrays are the raw slots of the reference's own data objects
(`data/seurat_sample_tms_liver.rda`, `data/LRI_mouse.rda`, `data/LRI_human.rda`; documented at
`R/data.R:1-175`). Pins asserted below come from SURVEY.md Appendix A.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rdx2  # noqa: E402

REF = os.environ.get("SCDC_REFERENCE", "/root/reference")
LRI_COLS = ["LRI", "LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3"]


def _str_array(values):
    """list of str/None -> numpy unicode array; NA_character_ -> '' (no gene symbol is empty)."""
    return np.array(["" if v is None else v for v in values], dtype=np.str_)


def liver():
    obj = rdx2.read_rda(f"{REF}/data/seurat_sample_tms_liver.rda")["seurat_sample_tms_liver"]
    mat = obj.attr["assays"]["RNA"].attr["data"]  # genes x cells dgCMatrix, log1p-normalised
    i = rdx2.unwrap(mat.attr["i"])
    p = rdx2.unwrap(mat.attr["p"])
    x = rdx2.unwrap(mat.attr["x"])
    dim = rdx2.unwrap(mat.attr["Dim"])
    dimnames = rdx2.unwrap(mat.attr["Dimnames"])
    genes = _str_array(rdx2.unwrap(dimnames[0]))
    cells = _str_array(rdx2.unwrap(dimnames[1]))
    md = obj.attr["meta.data"]
    cols = rdx2.dataframe_columns(md)
    row_names = _str_array(rdx2.unwrap(md.attr["row.names"]))
    assert tuple(dim) == (726, 468) and len(x) == 73215 and len(p) == 469
    assert (row_names == cells).all()  # R/utils_preprocessing.R:112-120
    out = dict(
        i=i.astype(np.int32), p=p.astype(np.int32), x=x.astype(np.float64),
        dim=dim.astype(np.int32), genes=genes, cells=cells,
        cell_type=_str_array(cols["cell_type"]), age_group=_str_array(cols["age_group"]),
    )
    np.savez_compressed(f"{HERE}/seurat_sample_tms_liver.npz", **out)
    print("liver:", {k: v.shape for k, v in out.items()})


def lri(species, n_expected):
    obj = rdx2.read_rda(f"{REF}/data/LRI_{species}.rda")[f"LRI_{species}"]
    curated = obj["LRI_curated"]
    cols = rdx2.dataframe_columns(curated)
    out = {c: _str_array(cols[c]) for c in LRI_COLS}
    assert len(out["LRI"]) == n_expected, len(out["LRI"])
    np.savez_compressed(f"{HERE}/LRI_{species}_curated.npz", **out)
    print(f"LRI_{species}:", len(out["LRI"]), "rows")


if __name__ == "__main__":
    liver()
    lri("mouse", 4582)
    lri("human", 4785)
