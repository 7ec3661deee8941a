"""CPU tests of the host-side mirror of run_stat_analysis (reference R/utils_permutation.R:1-115, 215-499): row subset,
observed vectors, gene subset, p = counts / iterations, merge-back with p = 1 outside the tested rows and NA for absent
subunits, distribution matrices. The device call is replaced by the CPU oracle here (no GPU in this container); the same
function runs against the engine in tests/test_parity_gpu.py::test_run_stat_analysis_mirror."""
import numpy as np
import pytest

from oracle import oracle_c, oracle_np
from scdiffcom_b200 import permutation
from tests import helpers


def _oracle_backed_perm_counts(prob, iterations, seed=42, score_type="geometric_mean", return_distributions=False, **kw):
    cond, ct = oracle_c.philox_labels(prob, int(iterations), seed=seed)
    counts, dist = oracle_c.perm_counts(prob, ct, cond, score_type=0 if score_type == "geometric_mean" else 1,
                                        return_distributions=return_distributions)
    return {"counts": counts, "distributions": dist, "stats": {}}


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_run_stat_analysis_host_logic(liver_cases, liver, mode, monkeypatch):
    monkeypatch.setattr(permutation, "perm_counts", _oracle_backed_perm_counts)
    case = liver_cases[mode]
    ai, simple, mask = case["ai"], case["simple"], case["mask"]
    res = permutation.run_stat_analysis(ai, simple, 20, True, "geometric_mean", False, seed=5)
    tab = res["cci_raw"]
    for k, v in simple.items():                                 # non-p-value columns unchanged (test-everything.R:534-561)
        assert np.array_equal(tab[k], v, equal_nan=True)
    names = permutation.pvalue_column_names(ai)
    assert names == (["P_VALUE"] if mode == "nocond" else
                     ["P_VALUE_DE", "P_VALUE_YOUNG", "P_VALUE_OLD", "L1_P_VALUE_DE", "L2_P_VALUE_DE", "R1_P_VALUE_DE",
                      "R2_P_VALUE_DE", "R3_P_VALUE_DE"])
    for n in names:
        col = tab[n]
        assert col.shape == mask.shape
        assert np.all((col[~np.isnan(col)] >= 0) & (col[~np.isnan(col)] <= 1))
        if n[0] in "LR" and n[1].isdigit():
            continue
        assert np.all(col[~mask] == 1.0) and not np.isnan(col).any()          # :434-441
    if mode == "cond":
        lig2_na = ai["LRI"]["LIGAND_2"][simple["LRI_IDX"]] == ""
        assert np.isnan(tab["L2_P_VALUE_DE"][lig2_na]).all()                   # :442-469
        assert not np.isnan(tab["L2_P_VALUE_DE"][~lig2_na]).any()
        assert np.all(tab["L2_P_VALUE_DE"][~lig2_na & ~mask] == 1.0)
    # p-values equal the NumPy oracle's on the same labels
    ai_o = helpers.liver_inputs(liver, mode, oracle_np)
    s_o = oracle_np.run_simple_cci_analysis(ai_o, oracle_np.create_cci_template(ai_o))
    prob = case["prob"]
    cond, ct = oracle_c.philox_labels(prob, 20, seed=5)
    cts = np.array(ai["cell_types"])
    a_sub, _ = oracle_np.restrict_to_tested(ai_o, s_o, oracle_np.tested_rows_and_observed(ai_o, s_o)[0])
    metas = []
    for q in range(20):
        if mode == "nocond":
            metas.append(({"cell_type": cts[ct[q]]},))
        else:
            cn = np.array(["YOUNG", "OLD"])[cond[q]]
            metas.append(({"cell_type": a_sub["metadata"]["cell_type"], "condition": cn}, {"cell_type": cts[ct[q]], "condition": cn}))
    r = oracle_np.run_stat_analysis(ai_o, s_o, 20, label_metas=metas)
    for k, n in enumerate(names):
        assert np.array_equal(tab[n][mask], r["pvals"][:, k], equal_nan=True), n
    # distributions: rows x (iterations + 1), last column = observed (:304-426)
    d = res["distributions"]
    key = "DISTRIBUTIONS" if mode == "nocond" else "DISTRIBUTIONS_DE"
    assert d[key].shape == (int(mask.sum()), 21)
    assert np.array_equal(d[key][:, -1], prob.observed.reshape(prob.n_rows, prob.ncols)[:, 0])


def test_iterations_above_1000_follow_the_reference_batching(liver_cases, monkeypatch):
    seen = {}

    def fake(prob, iterations, **kw):
        seen["n"] = iterations
        return {"counts": np.zeros((prob.n_rows, prob.ncols)), "distributions": None, "stats": {}}
    monkeypatch.setattr(permutation, "perm_counts", fake)
    case = liver_cases["nocond"]
    permutation.run_stat_analysis(case["ai"], case["simple"], 2500, False, "geometric_mean", False)
    assert seen["n"] == 1500          # R/utils_permutation.R:100-128: floor(2500/1000) batches, the last one 500
    permutation.run_stat_analysis(case["ai"], case["simple"], 3000, False, "geometric_mean", False)
    assert seen["n"] == 3000
