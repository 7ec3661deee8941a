"""SCDC_FP32_FAST (SURVEY.md section 8b options `precision`, `rel_tol`, `band_counts`; north_star: "exceedance counts bit-exact
... except for reported ties inside the tolerance band"): fp32 values and group sums, fp64 scores. Contract tested here:
every count that differs from the fp64 CPU oracle differs by at most the reported band count, the band is thin, and the
per-replicate scores are within 1e-5 relative of the oracle's fp64 scores (reference semantics being approximated:
R/utils_permutation.R:139-144,164-200)."""
import numpy as np
import pytest

from oracle import oracle_c
from scdiffcom_b200 import Engine, perm_counts, synthetic
from tests import helpers

pytestmark = pytest.mark.gpu
ST = {"geometric_mean": 0, "arithmetic_mean": 1}


def check_fp32_against_oracle(prob, n, seed, score_type="geometric_mean", with_dist=True, max_band_frac=0.02):
    pcond, pct = helpers.reference_like_labels(prob, n, seed)
    de_scale = np.zeros((n, prob.n_rows)) if (with_dist and prob.n_conditions == 2) else None
    want, want_d = oracle_c.perm_counts(prob, pct, pcond, score_type=ST[score_type], return_distributions=with_dist,
                                        de_scale=de_scale)
    got = perm_counts(prob, n, score_type=score_type, perm_cond=pcond, perm_ct=pct, precision="fp32_fast")
    c32, band = got["counts"], got["band_counts"]
    assert band is not None and band.shape == want.shape
    assert np.array_equal(np.isnan(c32), np.isnan(want)) and np.array_equal(np.isnan(band), np.isnan(want))
    diff = np.abs(np.nan_to_num(c32) - np.nan_to_num(want))
    assert np.all(diff <= np.nan_to_num(band)), f"{int((diff > np.nan_to_num(band)).sum())} counters differ by more than their band"
    assert np.nansum(band) <= max_band_frac * n * np.isfinite(want).sum(), "the tolerance band is not thin"
    if with_dist:
        gd = perm_counts(prob, n, score_type=score_type, perm_cond=pcond, perm_ct=pct, precision="fp32_fast",
                         return_distributions=True)
        # the generic kernel (distributions) and the production kernels agree on counts AND band membership
        assert np.array_equal(gd["counts"], c32, equal_nan=True)
        assert np.array_equal(gd["band_counts"], band, equal_nan=True)
        d32, d64 = gd["distributions"], want_d
        if prob.n_conditions == 1:
            assert np.all(np.abs(d32 - d64) <= 1e-5 * np.abs(d64))
        else:
            for k in (1, 2):   # CCI_SCORE of cond1 / cond2: within 1e-5 relative
                assert np.all(np.abs(d32[:, k] - d64[:, k]) <= 1e-5 * np.abs(d64[:, k])), k
            # DE = difference of two (pass-1) scores: its error scales with their sum, not with the difference
            assert np.all(np.abs(d32[:, 0] - d64[:, 0]) <= 1e-5 * de_scale)
    return got


@pytest.mark.parametrize("mode", ["cond", "nocond"])
@pytest.mark.parametrize("score_type", ["geometric_mean", "arithmetic_mean"])
def test_liver_fp32_fast_counts_differ_only_inside_the_band(liver, mode, score_type):
    case = helpers.liver_case(liver, mode, score_type)
    got = check_fp32_against_oracle(case["prob"], 200, seed=5, score_type=score_type)
    assert got["stats"]["kernel_launches"] > 0


@pytest.mark.parametrize("kw", [
    dict(n_cells=3000, n_genes=200, n_celltypes=6, n_conditions=2, n_lri=300, seed=1),     # K=12  -> J=4
    dict(n_cells=9000, n_genes=120, n_celltypes=30, n_conditions=2, n_lri=200, seed=3),    # K=60  -> J=2
    dict(n_cells=2500, n_genes=180, n_celltypes=9, n_conditions=1, n_lri=300, seed=4),     # detection, K=9
    dict(n_cells=20000, n_genes=60, n_celltypes=60, n_conditions=2, n_lri=100, seed=5, min_group=20),  # K=120 -> J=1
])
def test_synthetic_fp32_fast(kw):
    prob, ai, simple, mask = synthetic.make_problem(kw)
    check_fp32_against_oracle(prob, 160, seed=kw["seed"])


def test_band_grows_with_rel_tol_and_fp64_reports_none(liver_cases):
    prob = liver_cases["cond"]["prob"]
    pcond, pct = helpers.reference_like_labels(prob, 128, 9)
    kw = dict(perm_cond=pcond, perm_ct=pct, precision="fp32_fast")
    dflt = perm_counts(prob, 128, **kw)
    same = perm_counts(prob, 128, rel_tol=1e-5, **kw)
    wide = perm_counts(prob, 128, rel_tol=1e-3, **kw)
    assert np.array_equal(dflt["band_counts"], same["band_counts"], equal_nan=True)       # rel_tol <= 0 means 1e-5
    assert np.array_equal(dflt["counts"], wide["counts"], equal_nan=True)                 # the band never changes a count
    assert np.all(np.nan_to_num(wide["band_counts"]) >= np.nan_to_num(dflt["band_counts"]))
    assert np.nansum(wide["band_counts"]) > np.nansum(dflt["band_counts"])
    exact = perm_counts(prob, 128, perm_cond=pcond, perm_ct=pct)
    assert exact["band_counts"] is None
    # device-drawn labels, resident engine, sharded replicate ranges: counts and band counters are additive
    with Engine(prob) as eng:
        whole = eng.run(500, seed=3, precision="fp32_fast")
        a = eng.run(200, seed=3, precision="fp32_fast")
        b = eng.run(300, seed=3, perm_offset=200, precision="fp32_fast")
        again64 = eng.run(500, seed=3)          # the same engine serves both precisions
    for key in ("counts", "band_counts"):
        assert np.array_equal(np.nan_to_num(whole[key]), np.nan_to_num(a[key]) + np.nan_to_num(b[key])), key
    cond_o, ct_o = oracle_c.philox_labels(prob, 500, seed=3)
    want, _ = oracle_c.perm_counts(prob, ct_o, cond_o)
    assert np.array_equal(again64["counts"], want, equal_nan=True)
    assert np.all(np.abs(np.nan_to_num(whole["counts"]) - np.nan_to_num(want)) <= np.nan_to_num(whole["band_counts"]))
