"""SURVEY.md "next" item f3 (CPU tests, no device needed): the integer-coded CJ template + NCELLS (`scdc_cci_template`, reference
create_cci_template / add_cell_number, R/utils_cci.R:1-169) and the merge-back of the p-values (`scdc_merge_pvalues`,
R/utils_permutation.R:215-259, 427-469), both through the C-ABI, against the string-keyed NumPy oracle."""
import time

import numpy as np
import pytest

from oracle import oracle_np
from scdiffcom_b200 import _lib, cci, engine, synthetic
from tests import helpers


def _decode(ai, tmpl):
    cts = np.array(ai["cell_types"])
    out = {"EMITTER_CELLTYPE": cts[tmpl["EMITTER_CODE"]], "RECEIVER_CELLTYPE": cts[tmpl["RECEIVER_CODE"]],
           "LRI": np.asarray(ai["LRI"]["LRI"])[tmpl["LRI_IDX"]]}
    for c in ai["LRI"]:
        if c != "LRI":
            out[c] = np.asarray(ai["LRI"][c])[tmpl["LRI_IDX"]]
    out.update({k: v for k, v in tmpl.items() if k.startswith("NCELLS")})
    return out


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_template_from_the_shim_equals_the_oracle_on_the_liver_fixture(liver, mode):
    from scdiffcom_b200 import preprocessing
    ai = helpers.liver_inputs(liver, mode, preprocessing)
    got = _decode(ai, cci.create_cci_template(ai))
    want = oracle_np.create_cci_template(helpers.liver_inputs(liver, mode, oracle_np))
    assert len(got["LRI"]) == 29_325                        # tests/testthat/test-everything.R:281-294
    assert set(got) == set(want)
    for k, v in want.items():
        assert np.array_equal(got[k], v), k


def test_template_and_merge_back_on_a_synthetic_atlas_shape():
    """T = 60 cell types, differential mode, LRI_human-like complexes: template vs the oracle's cross join, then the
    merge-back of synthetic counts vs a plain NumPy restatement of R/utils_permutation.R:427-469."""
    ai = synthetic.make_analysis_inputs(n_cells=9000, n_genes=120, n_celltypes=60, n_conditions=2, n_lri=300, seed=77,
                                        min_group=10)
    tmpl = cci.create_cci_template(ai)
    got = _decode(ai, tmpl)
    want = oracle_np.create_cci_template(ai)
    for k, v in want.items():
        assert np.array_equal(got[k], v), k
    n_full = len(tmpl["LRI_IDX"])
    assert n_full == 60 * 60 * 300
    rng = np.random.default_rng(3)
    mask = rng.random(n_full) < 0.3
    rows = np.flatnonzero(mask)
    enc = cci.encode_inputs(ai)
    nL, nR = ai["max_nL"], ai["max_nR"]
    ncols = 3 + nL + nR
    counts = rng.integers(0, 1001, size=(len(rows), ncols)).astype(np.uint64)
    absent = enc["sub_gene"][tmpl["LRI_IDX"]] < 0                       # [n_full, nS]
    counts[:, 3:][absent[rows]] = np.uint64(_lib.SCDC_COUNT_NA)         # the engine marks NA subunits this way
    pv = engine.merge_pvalues(n_full, rows, counts, 1000, sub_gene=enc["sub_gene"], lri_idx=tmpl["LRI_IDX"], max_nL=nL, max_nR=nR)
    want_pv = np.ones((n_full, ncols))                                   # :434-441
    want_pv[rows] = counts.astype(np.float64) / 1000                     # :215-259
    want_pv[:, 3:][absent] = np.nan                                      # :442-469
    assert np.array_equal(pv, want_pv, equal_nan=True)
    # detection mode: one column, no subunit handling
    pv1 = engine.merge_pvalues(n_full, rows, counts[:, :1], 1000)
    assert np.array_equal(pv1[:, 0], want_pv[:, 0])


def test_full_atlas_template_size_is_seconds_not_minutes():
    """BASELINE.json config 5: T = 60, 5000 LRIs -> 18 M rows x (3 + 4) integer columns, and the merge-back of 8 p-value
    columns over 18 M rows with 5.8 M tested: the R-side CJ + string merges this replaces dwarf a 3 s permutation run."""
    T, L, N = 60, 5000, 250_000
    rng = np.random.default_rng(5)
    ct = rng.integers(0, T, size=N).astype(np.int32)
    cd = rng.integers(0, 2, size=N).astype(np.int32)
    order = rng.permutation(L).astype(np.int32)
    t0 = time.perf_counter()
    em, rc, li, ne, nr = engine.cci_template(T, 2, order, ct, cd)
    t_tmpl = time.perf_counter() - t0
    n = T * T * L
    assert em.shape == (n,) and ne.shape == (2, n)
    i = rng.integers(0, n, size=1000)
    assert np.array_equal(em[i], i // (T * L)) and np.array_equal(rc[i], (i // L) % T) and np.array_equal(li[i], order[i % L])
    nc = np.zeros((2, T), np.int64)
    np.add.at(nc, (cd, ct), 1)
    assert np.array_equal(ne[:, i], nc[:, em[i]]) and np.array_equal(nr[:, i], nc[:, rc[i]])
    rows = np.flatnonzero(rng.random(n) < 0.32)
    counts = rng.integers(0, 100_000, size=(len(rows), 8)).astype(np.uint64)
    sub_gene = np.where(rng.random((L, 5)) < 0.7, rng.integers(0, 2000, size=(L, 5)), -1).astype(np.int32)
    sub_gene[:, [0, 2]] = np.abs(sub_gene[:, [0, 2]])
    t0 = time.perf_counter()
    pv = engine.merge_pvalues(n, rows, counts, 100_000, sub_gene=sub_gene, lri_idx=li, max_nL=2, max_nR=3)
    t_merge = time.perf_counter() - t0
    assert pv.shape == (n, 8) and np.all(pv[rows[:100], 0] == counts[:100, 0] / 100_000)
    untested = np.setdiff1d(np.arange(0, n, 9973), rows)
    assert np.all(pv[untested, :3] == 1.0)
    assert np.array_equal(np.isnan(pv[untested, 3:]), sub_gene[li[untested]] < 0)
    assert t_tmpl < 20 and t_merge < 30, (t_tmpl, t_merge)


def test_argument_errors():
    with pytest.raises(_lib.ScdcError, match="permutation"):
        engine.cci_template(3, 1, np.array([0, 0, 1]), np.zeros(5, np.int32))
    with pytest.raises(_lib.ScdcError, match="out of range"):
        engine.merge_pvalues(10, np.array([11]), np.zeros((1, 1), np.uint64), 10)


def _p_adjust_bh(p):
    """stats::p.adjust(p, "BH") restated in NumPy: NA left out, pmin(1, cummin(n / i * p[o]))[ro] over decreasing p."""
    p = np.asarray(p, dtype=np.float64)
    out = np.full_like(p, np.nan)
    ok = ~np.isnan(p)
    q = p[ok]
    n = len(q)
    if n:
        o = np.argsort(-q, kind="stable")
        adj = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * q[o]))
        r = np.empty(n)
        r[o] = adj
        out[ok] = r
    return out


def test_bh_adjustment_matches_p_adjust():
    """f4: the BH_P_VALUE* columns of process_cci_raw (R/utils_filtering.R:216-235)."""
    rng = np.random.default_rng(11)
    p = np.round(rng.random(50_000), 3)                # many ties, like counts / iterations
    p[rng.random(p.size) < 0.1] = np.nan               # NA subunits
    p[:5] = [0.0, 1.0, 1.0, 0.0, np.nan]
    got = engine.bh_adjust(p)
    assert np.array_equal(got, _p_adjust_bh(p), equal_nan=True)
    from scipy.stats import false_discovery_control    # an independent implementation (no NA support)
    ok = ~np.isnan(p)
    np.testing.assert_allclose(got[ok], false_discovery_control(p[ok], method="bh"), rtol=1e-14)
    g = rng.integers(0, 7, size=p.size).astype(np.int32)
    got_g = engine.bh_adjust(p, g)
    for k in range(7):
        assert np.array_equal(got_g[g == k], _p_adjust_bh(p[g == k]), equal_nan=True)
    assert engine.bh_adjust(np.array([])).size == 0
