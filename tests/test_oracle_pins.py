"""CPU tests: the oracle against the reference's own fixtures, golden numbers and known-answer identities.

Pins (SURVEY.md sections 4 and 8c): reference tests/testthat/test-everything.R:281-294 (29 325 template rows), :377-490
(mean / mean(x>0) / sqrt(mean*mean) identities, tolerance 1.5e-8), docs/articles/scDiffCom-vignette.html:155-161,
295-301, 391-397 (8952 / 7352 / 111 tested CCIs), data documented at R/data.R.
"""
import numpy as np
import pytest

from oracle import oracle_c, oracle_np
from tests import helpers


@pytest.fixture(scope="module")
def oracle_cases(liver):
    out = {}
    for mode in ("cond", "nocond"):
        ai = helpers.liver_inputs(liver, mode, oracle_np)
        tmpl = oracle_np.create_cci_template(ai)
        out[mode] = (ai, oracle_np.run_simple_cci_analysis(ai, tmpl))
    return out


def test_fixture_shape_pins(liver):
    assert liver["X"].shape == (726, 468) and liver["X"].nnz == 73215
    assert len(liver["LRI_mouse"]["LRI"]) == 4582
    ct, n = np.unique(liver["cell_type"], return_counts=True)
    assert len(ct) == 5
    tab = {(c, a): int(np.sum((liver["cell_type"] == c) & (liver["age_group"] == a))) for c in ct for a in ("YOUNG", "OLD")}
    assert sorted(tab.values()) == [34, 34] + [50] * 8
    assert tab[("T cell", "YOUNG")] == 34 and tab[("B cell", "YOUNG")] == 34


def test_extraction_pins(oracle_cases):
    ai, _ = oracle_cases["cond"]
    assert ai["data_tr"].shape == (468, 652) and ai["data_tr"].nnz == 62298
    assert len(ai["LRI"]["LRI"]) == 1173 and ai["max_nL"] == 2 and ai["max_nR"] == 3
    assert int(np.sum(ai["LRI"]["LIGAND_2"] != "")) == 13
    assert int(np.sum(ai["LRI"]["RECEPTOR_2"] != "")) == 184 and int(np.sum(ai["LRI"]["RECEPTOR_3"] != "")) == 6


def test_template_and_tested_row_pins(oracle_cases):
    for mode, n_tested in (("cond", 8952), ("nocond", 7352)):
        ai, simple = oracle_cases[mode]
        assert len(simple["LRI"]) == 29325                     # test-everything.R:281-294
        mask, obs = oracle_np.tested_rows_and_observed(ai, simple)
        assert int(mask.sum()) == n_tested                     # vignette html:155-161, 295-301
        assert obs.shape == (n_tested, 8 if mode == "cond" else 1)


def test_custom_100_lri_pin(liver):
    lri = {k: v[:100] for k, v in liver["LRI_mouse"].items()}   # vignette: first 100 curated LRIs
    ai = helpers.liver_inputs(liver, "cond", oracle_np, lri=lri)
    simple = oracle_np.run_simple_cci_analysis(ai, oracle_np.create_cci_template(ai))
    mask, _ = oracle_np.tested_rows_and_observed(ai, simple)
    assert len(ai["LRI"]["LRI"]) == 16 and len(simple["LRI"]) == 400 and int(mask.sum()) == 111   # html:391-397


def test_known_answer_identities(liver, oracle_cases):
    """test-everything.R:296-490: hepatocyte -> endothelial cell of hepatic sinusoid, Adam15:Itga5."""
    X = liver["X"].tocsr()
    genes = liver["genes"].tolist()
    e = np.expm1(X[genes.index("Adam15")].toarray().ravel())
    r = np.expm1(X[genes.index("Itga5")].toarray().ravel())
    em, rc = "hepatocyte", "endothelial cell of hepatic sinusoid"
    ct, age = liver["cell_type"], liver["age_group"]

    def row(simple):
        m = (simple["EMITTER_CELLTYPE"] == em) & (simple["RECEIVER_CELLTYPE"] == rc) & (simple["LRI"] == "Adam15:Itga5")
        assert m.sum() == 1
        return {k: v[m][0] for k, v in simple.items()}

    tol = dict(rtol=1.5e-8, atol=0)
    _, s = oracle_cases["nocond"]
    x = row(s)
    np.testing.assert_allclose(x["L1_DETECTION_RATE"], np.mean(e[ct == em] > 0), **tol)
    np.testing.assert_allclose(x["R1_DETECTION_RATE"], np.mean(r[ct == rc] > 0), **tol)
    np.testing.assert_allclose(x["L1_EXPRESSION"], np.mean(e[ct == em]), **tol)
    np.testing.assert_allclose(x["CCI_SCORE"], np.sqrt(np.mean(e[ct == em]) * np.mean(r[ct == rc])), **tol)
    assert x["CCI_SCORE"] == 0.009375777142019296
    _, s = oracle_cases["cond"]
    x = row(s)
    for a in ("YOUNG", "OLD"):
        ea, ra = e[(ct == em) & (age == a)], r[(ct == rc) & (age == a)]
        np.testing.assert_allclose(x[f"L1_DETECTION_RATE_{a}"], np.mean(ea > 0), **tol)
        np.testing.assert_allclose(x[f"R1_EXPRESSION_{a}"], np.mean(ra), **tol)
        np.testing.assert_allclose(x[f"CCI_SCORE_{a}"], np.sqrt(np.mean(ea) * np.mean(ra)), **tol)


def test_host_mirror_matches_oracle_observed_pass(liver_cases, oracle_cases):
    for mode in ("cond", "nocond"):
        ai, s = liver_cases[mode]["ai"], liver_cases[mode]["simple"]
        ai_o, s_o = oracle_cases[mode]
        names = np.array(ai["cell_types"])
        assert (names[s["EMITTER_CODE"]] == s_o["EMITTER_CELLTYPE"]).all()
        assert (names[s["RECEIVER_CODE"]] == s_o["RECEIVER_CELLTYPE"]).all()
        assert (ai["LRI"]["LRI"][s["LRI_IDX"]] == s_o["LRI"]).all()
        checked = 0
        for k, v in s.items():
            if k in s_o and v.dtype.kind in "fb":
                assert np.array_equal(v, s_o[k], equal_nan=True), k
                checked += 1
        assert checked >= (9 if mode == "nocond" else 20)
        assert liver_cases[mode]["prob"].n_genes == (645 if mode == "cond" else 612)
        if mode == "cond":   # per-subunit LOGFC columns (R/utils_cci.R:284-437): NA exactly where the subunit is absent
            for name, col in (("L1", "LIGAND_1"), ("L2", "LIGAND_2"), ("R1", "RECEPTOR_1"), ("R3", "RECEPTOR_3")):
                v = s[f"{name}_LOGFC"]
                assert np.array_equal(v, s_o[f"{name}_LOGFC"], equal_nan=True)
                assert np.array_equal(np.isnan(v), s_o[col] == "")


@pytest.mark.parametrize("mode", ["cond", "nocond"])
@pytest.mark.parametrize("score_type", ["geometric_mean", "arithmetic_mean"])
def test_c_oracle_equals_numpy_oracle(liver, mode, score_type):
    """The two independent restatements agree bit for bit (counts and per-replicate values) on the same labels."""
    ai = helpers.liver_inputs(liver, mode, oracle_np)
    simple = oracle_np.run_simple_cci_analysis(ai, oracle_np.create_cci_template(ai), score_type=score_type)
    case = helpers.liver_case(liver, mode, score_type)
    prob = case["prob"]
    n = 6
    pcond, pct = helpers.reference_like_labels(prob, n, seed=11)
    names = np.array(ai["cell_types"])
    metas = []
    a_sub, _ = oracle_np.restrict_to_tested(ai, simple, oracle_np.tested_rows_and_observed(ai, simple)[0])
    for q in range(n):
        if mode == "nocond":
            metas.append(({"cell_type": names[pct[q]]},))
        else:
            cn = np.array(["YOUNG", "OLD"])[pcond[q]]
            metas.append(({"cell_type": a_sub["metadata"]["cell_type"], "condition": cn},
                          {"cell_type": names[pct[q]], "condition": cn}))
    r = oracle_np.run_stat_analysis(ai, simple, n, score_type=score_type, label_metas=metas, return_distributions=True)
    cc, dist = oracle_c.perm_counts(prob, pct, pcond, score_type=0 if score_type == "geometric_mean" else 1,
                                    return_distributions=True)
    expect = np.where(np.isnan(r["observed"]), np.nan, r["counts"].astype(float))
    assert np.array_equal(cc, expect, equal_nan=True)
    d_np = np.moveaxis(r["distributions"], 2, 0)          # [it, R, ncols]
    ncd = 3 if mode == "cond" else 1
    assert np.array_equal(dist, np.moveaxis(d_np[:, :, :ncd], 2, 1))


def test_batching_quirk():
    """R/utils_permutation.R:100-128."""
    from scdiffcom_b200.permutation import iterations_to_run
    for f in (oracle_np.iterations_actually_run, iterations_to_run):
        assert f(10) == 10 and f(1000) == 1000 and f(2000) == 2000 and f(10000) == 10000
        assert f(2500) == 1500 and f(1001) == 1 and f(1999) == 999
    with pytest.raises(ValueError):
        iterations_to_run(1001, return_distributions=True)


def test_oracle_philox_labels_are_stratified_shuffles(liver_cases):
    for mode in ("cond", "nocond"):
        prob = liver_cases[mode]["prob"]
        cond, ct = oracle_c.philox_labels(prob, 20, seed=3)
        T = prob.n_celltypes
        for q in range(20):
            if mode == "cond":
                base = np.bincount(np.asarray(prob.cond_code) * T + np.asarray(prob.ct_code), minlength=2 * T)
                assert np.array_equal(np.bincount(cond[q].astype(int) * T + np.asarray(prob.ct_code), minlength=2 * T), base)
                assert np.array_equal(np.bincount(cond[q].astype(int) * T + ct[q], minlength=2 * T), base)
            else:
                assert np.array_equal(np.bincount(ct[q], minlength=T), np.bincount(prob.ct_code, minlength=T))
        # counter-based: replicate q does not depend on where the range starts
        cond2, ct2 = oracle_c.philox_labels(prob, 5, seed=3, perm_offset=7)
        assert np.array_equal(ct2, ct[7:12])


@pytest.mark.parametrize("urn", ["0", "1"])
def test_oracle_label_generators_are_uniform_within_strata(liver_cases, urn, monkeypatch):
    """Both restated generators (random keys, urn) draw every cell's permuted condition with probability
    n(t, cond2) / n_t inside its cell type, and its permuted cell type from the mixture over the permuted condition
    (the distribution of base::sample on the reference's strata, R/utils_permutation.R:536-569)."""
    monkeypatch.setenv("SCDC_URN_LABELS", urn)
    prob = liver_cases["cond"]["prob"]
    n = 2000
    cond, ct = oracle_c.philox_labels(prob, n, seed=11)
    T, N = prob.n_celltypes, prob.n_cells
    ct0, cd0 = np.asarray(prob.ct_code), np.asarray(prob.cond_code)
    base = np.zeros((2, T), np.int64)
    np.add.at(base, (cd0, ct0), 1)
    p1 = (base[1] / base.sum(0))[ct0]
    z = (cond.mean(0) - p1) / np.sqrt(p1 * (1 - p1) / n)
    assert np.abs(z).max() < 5.5 and abs((z ** 2).sum() - N) < 6 * np.sqrt(2 * N)
    pc = np.stack([1 - p1, p1], axis=1)
    pt = pc @ (base / base.sum(1, keepdims=True))
    obs = np.stack([(ct == t).sum(0) for t in range(T)], axis=1)
    chi2 = ((obs - n * pt) ** 2 / (n * pt)).sum()
    dof = N * (T - 1)
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)


def _naive_counts(prob, pcond, pct, score_type):
    """Third, deliberately naive restatement for tiny problems (pure Python loops, dense means): follows
    run_stat_iteration -> aggregate_cells -> build_cci_or_drate -> counting literally, one replicate at a time."""
    import math
    N, G, T, C = prob.n_cells, prob.n_genes, prob.n_celltypes, prob.n_conditions
    nL, nR = prob.max_nL, prob.max_nR
    colptr, rowidx, values = np.asarray(prob.colptr), np.asarray(prob.rowidx), np.asarray(prob.values)
    obs = np.asarray(prob.observed).reshape(prob.n_rows, prob.ncols)
    sub = np.asarray(prob.sub_gene).reshape(-1, nL + nR)
    counts = np.zeros(obs.shape, dtype=np.int64)

    def means(group):
        K = C * T
        m = [[0.0] * G for _ in range(K)]
        n = [0] * K
        for i in range(N):
            n[group[i]] += 1
        for g in range(G):
            for e in range(colptr[g], colptr[g + 1]):
                m[group[rowidx[e]]][g] += float(values[e])      # ascending cell order inside the column
        return [[m[k][g] / n[k] for g in range(G)] for k in range(K)]

    def score(m, c, row):
        l, em, rc = prob.row_lri[row], prob.row_emit[row], prob.row_recv[row]
        ex = [float("nan") if sub[l, s] < 0 else m[c * T + (em if s < nL else rc)][sub[l, s]] for s in range(nL + nR)]
        mL = min(v for v in ex[:nL] if not math.isnan(v))
        mR = min(v for v in ex[nL:] if not math.isnan(v))
        return (math.sqrt(mL * mR) if score_type == 0 else (mL + mR) / 2), ex

    for q in range(pct.shape[0]):
        if C == 1:
            m2 = means([int(x) for x in pct[q]])
            for r in range(prob.n_rows):
                counts[r, 0] += score(m2, 0, r)[0] >= obs[r, 0]
        else:
            m1 = means([int(pcond[q, i]) * T + int(prob.ct_code[i]) for i in range(N)])
            m2 = means([int(pcond[q, i]) * T + int(pct[q, i]) for i in range(N)])
            for r in range(prob.n_rows):
                (a1, e1), (a2, e2) = score(m1, 0, r), score(m1, 1, r)
                counts[r, 0] += abs(a2 - a1) >= abs(obs[r, 0])
                counts[r, 1] += score(m2, 0, r)[0] >= obs[r, 1]
                counts[r, 2] += score(m2, 1, r)[0] >= obs[r, 2]
                for s in range(nL + nR):
                    d = e2[s] - e1[s]
                    counts[r, 3 + s] += (not math.isnan(d)) and (not math.isnan(obs[r, 3 + s])) and abs(d) >= abs(obs[r, 3 + s])
    out = counts.astype(float)
    out[np.isnan(obs)] = np.nan
    return out


@pytest.mark.parametrize("kw", [
    dict(seed=21, C=1, N=40, G=4, T=3, L=3, nL=1, nR=2, R=12),
    dict(seed=22, C=2, N=60, G=5, T=2, L=4, nL=2, nR=3, R=25),
    dict(seed=23, C=2, N=90, G=6, T=4, L=5, nL=2, nR=2, R=40),
])
@pytest.mark.parametrize("score_type", [0, 1])
def test_c_oracle_equals_naive_python_restatement(kw, score_type):
    """Tiny random problems with empty gene columns, absent subunits, exact ties between a replicate and the observed
    values (the `>=` of reference R/utils_permutation.R:141-143 counts them), zero and sign-flipped observed values."""
    prob, rng = helpers.random_problem(**kw)
    pcond, pct = helpers.with_tying_observed(prob, rng, 12)
    want = _naive_counts(prob, pcond, pct, score_type)
    got, _ = oracle_c.perm_counts(prob, pct, pcond, score_type=score_type)
    assert np.array_equal(got, want, equal_nan=True)
    if score_type == 0:
        assert (got[:, 0] >= 1).all()      # replicate 0 ties with the observed value and is counted
