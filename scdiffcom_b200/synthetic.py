"""Synthetic `analysis_inputs` of the shapes BASELINE.json names (SURVEY.md section 8d "Synthetic inputs").

Seed = 20261019 + config number. Matrix cells x genes CSC with int32 indices; per-gene density ~ Beta(0.8, 3.2) clipped
to [0.01, 0.99] (mean 0.20), modulated per (gene, cell type) by LogNormal(0, 0.5); non-zero values
exp(Normal(0.2, 2.4)) x per-(gene, cell type) LogNormal(0, 1) effect, and for 10 % of (gene, cell type) a x2 / x0.5
condition effect in cond2. Cell-type sizes ~ Dirichlet(2) with a floor per (cell type, condition); condition
Bernoulli(0.5). LRI table: `n_lri` synthetic LRIs over the G genes with LRI_human's (nL, nR) histogram.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# (nL, nR) histogram of LRI_human$LRI_curated
LRI_HUMAN_HIST = {(1, 1): 3648, (1, 2): 1096, (2, 2): 22, (2, 1): 9, (1, 3): 7, (2, 3): 3}

CONFIGS = {
    # name: N cells, G genes, T cell types, C conditions, L LRIs, permutations (BASELINE.json configs[1..4])
    "config2": dict(n_cells=20_000, n_genes=2500, n_celltypes=15, n_conditions=1, n_lri=5000, iterations=10_000, seed=20261021),
    "config3": dict(n_cells=50_000, n_genes=2500, n_celltypes=25, n_conditions=2, n_lri=5000, iterations=10_000, seed=20261022),
    "config4": dict(n_cells=100_000, n_genes=1854, n_celltypes=30, n_conditions=2, n_lri=4785, iterations=50_000, seed=20261023),
    # scaled-down stand-in for config 5 (same T = 60, K = 120 kernel regime) for quick tuning runs
    "config5s": dict(n_cells=60_000, n_genes=400, n_celltypes=60, n_conditions=2, n_lri=800, iterations=4096, seed=20261025),
    "config5": dict(n_cells=250_000, n_genes=2500, n_celltypes=60, n_conditions=2, n_lri=5000, iterations=100_000, seed=20261024),
}


def synthetic_lri_table(n_lri, genes, rng):
    keys = list(LRI_HUMAN_HIST)
    prob = np.array([LRI_HUMAN_HIST[k] for k in keys], dtype=np.float64)
    kind = rng.choice(len(keys), size=n_lri, p=prob / prob.sum())
    G = len(genes)
    cols = {c: np.full(n_lri, "", dtype=genes.dtype) for c in ("LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3")}
    for k, (nl, nr) in enumerate(keys):
        idx = np.flatnonzero(kind == k)
        if len(idx) == 0:
            continue
        # distinct genes inside one LRI
        pick = np.argsort(rng.random((len(idx), G)), axis=1)[:, :nl + nr] if G <= 64 else \
            np.stack([rng.choice(G, size=nl + nr, replace=False) for _ in range(len(idx))])
        for s in range(nl):
            cols[f"LIGAND_{s + 1}"][idx] = genes[pick[:, s]]
        for s in range(nr):
            cols[f"RECEPTOR_{s + 1}"][idx] = genes[pick[:, nl + s]]
    table = {"LRI": np.array([f"LRI{k:05d}" for k in range(n_lri)])}
    table.update(cols)
    return table


def make_analysis_inputs(n_cells, n_genes, n_celltypes, n_conditions, n_lri, seed, min_group=50, iterations=None,
                         density_scale=1.0, lri_table=None):
    """Returns the `analysis_inputs` dict (same keys as preprocessing.extract_analysis_inputs).
    lri_table: optional real LRI table (dict with LRI, LIGAND_1..2, RECEPTOR_1..3 columns, "" = NA); the matrix genes are
    then named after its gene symbols (BASELINE.json config 4: LRI_human) and n_genes / n_lri are taken from it."""
    rng = np.random.default_rng(seed)
    real_genes = None
    if lri_table is not None:
        cols = ("LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3")
        real_genes = np.unique(np.concatenate([np.asarray(lri_table[c]) for c in cols]))
        real_genes = real_genes[real_genes != ""]
        n_genes = len(real_genes)
    N, G, T, C = n_cells, n_genes, n_celltypes, n_conditions
    floor = min_group * C + 20 * C
    if floor * T > N:
        raise ValueError("too few cells for the per-group floor")
    sizes = np.full(T, floor) + rng.multinomial(N - floor * T, rng.dirichlet(np.full(T, 2.0)))
    ct_code = rng.permutation(np.repeat(np.arange(T), sizes)).astype(np.int32)
    cond_code = None
    if C == 2:
        cond_code = (rng.random(N) < 0.5).astype(np.int32)
        for t in range(T):            # enforce the floor per (cell type, condition)
            for c in (0, 1):
                idx = np.flatnonzero((ct_code == t) & (cond_code == c))
                if len(idx) < min_group:
                    other = np.flatnonzero((ct_code == t) & (cond_code == 1 - c))
                    cond_code[other[:min_group - len(idx)]] = c
    dens = np.clip(rng.beta(0.8, 3.2, size=G) * density_scale, 0.01, 0.99)
    m = np.exp(rng.normal(0.0, 0.5, size=(G, T)))
    eff = np.exp(rng.normal(0.0, 1.0, size=(G, T)))
    cond_eff = np.ones((G, T))
    if C == 2:
        hit = rng.random((G, T)) < 0.10
        cond_eff[hit] = rng.choice([2.0, 0.5], size=int(hit.sum()))
    indptr = np.zeros(G + 1, dtype=np.int64)
    idx_chunks, val_chunks = [], []
    for g in range(G):
        p = np.minimum(dens[g] * m[g, ct_code], 0.99)
        nz = np.flatnonzero(rng.random(N, dtype=np.float32) < p).astype(np.int32)
        v = np.exp(rng.normal(0.2, 2.4, size=len(nz))) * eff[g, ct_code[nz]]
        if C == 2:
            v = v * np.where(cond_code[nz] == 1, cond_eff[g, ct_code[nz]], 1.0)
        idx_chunks.append(nz)
        val_chunks.append(v)
        indptr[g + 1] = indptr[g] + len(nz)
    if indptr[-1] >= 2**31:
        raise ValueError("nnz exceeds int32 (dgCMatrix limit)")
    data_tr = sp.csc_matrix((np.concatenate(val_chunks), np.concatenate(idx_chunks), indptr.astype(np.int32)), shape=(N, G))
    if real_genes is not None:
        genes = real_genes
        table = np.stack([np.asarray(lri_table[c], dtype=np.str_) for c in ("LRI", "LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3")], axis=1)
        _, first = np.unique(table, axis=0, return_index=True)
        table = table[np.sort(first)]
        LRI = {c: table[:, k] for k, c in enumerate(("LRI", "LIGAND_1", "LIGAND_2", "RECEPTOR_1", "RECEPTOR_2", "RECEPTOR_3"))}
    else:
        genes = np.array([f"G{g:05d}" for g in range(G)])
        LRI = synthetic_lri_table(n_lri, genes, rng)
    max_nL = 2 if np.any(LRI["LIGAND_2"] != "") else 1
    max_nR = 3 if np.any(LRI["RECEPTOR_3"] != "") else (2 if np.any(LRI["RECEPTOR_2"] != "") else 1)
    for c in ["LIGAND_2"] * (max_nL < 2) + ["RECEPTOR_3"] * (max_nR < 3) + ["RECEPTOR_2"] * (max_nR < 2):
        LRI.pop(c)
    cell_types = [f"CT{t:03d}" for t in range(T)]
    metadata = {"cell_id": np.array([f"C{i:07d}" for i in range(N)]), "cell_type": np.array(cell_types)[ct_code]}
    cond = {"is_cond": C == 2, "is_samp": False}
    if C == 2:
        metadata["condition"] = np.array(["YOUNG", "OLD"])[cond_code]
        cond.update(cond1="YOUNG", cond2="OLD")
    return {"data_tr": data_tr, "genes": genes, "metadata": metadata, "cell_types": cell_types, "LRI": LRI,
            "max_nL": max_nL, "max_nR": max_nR, "condition": cond}


def make_problem(name_or_kwargs, score_type="geometric_mean", lri_table=None):
    """Synthetic inputs -> observed pass -> (PermutationProblem, analysis_inputs, cci_dt_simple, tested mask)."""
    from . import cci, permutation
    kw = dict(CONFIGS[name_or_kwargs]) if isinstance(name_or_kwargs, str) else dict(name_or_kwargs)
    kw.pop("iterations", None)
    ai = make_analysis_inputs(lri_table=lri_table, **kw)
    tmpl = cci.create_cci_template(ai)
    simple = cci.run_simple_cci_analysis(ai, tmpl, score_type=score_type)
    prob, mask = permutation.build_problem(ai, simple)
    return prob, ai, simple, mask
