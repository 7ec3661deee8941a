"""GPU parity tests (run with `pytest -m gpu` on a B200): the CUDA path, called through the C-ABI, against the CPU oracle
on the same inputs. Integer counts must be bit-exact; per-replicate fp64 values must be bit-exact too (same summation
order, IEEE /, *, sqrt), which is stricter than north_star's 1e-5 relative tolerance."""
import numpy as np
import pytest

from oracle import oracle_c
from scdiffcom_b200 import Engine, perm_counts, run_stat_analysis, synthetic
from scdiffcom_b200 import _lib
from tests import helpers

pytestmark = pytest.mark.gpu
ST = {"geometric_mean": 0, "arithmetic_mean": 1}


def _check_uploaded(prob, n, seed, score_type="geometric_mean", batch_size=0, dist=False):
    pcond, pct = helpers.reference_like_labels(prob, n, seed)
    want, want_d = oracle_c.perm_counts(prob, pct, pcond, score_type=ST[score_type], return_distributions=dist)
    got = perm_counts(prob, n, score_type=score_type, perm_cond=pcond, perm_ct=pct, batch_size=batch_size,
                      return_distributions=dist)
    assert np.array_equal(got["counts"], want, equal_nan=True)
    if dist:
        assert np.array_equal(got["distributions"], want_d)
    return got


@pytest.mark.parametrize("mode", ["cond", "nocond"])
@pytest.mark.parametrize("score_type", ["geometric_mean", "arithmetic_mean"])
def test_liver_uploaded_labels_bit_exact(liver, mode, score_type):
    """BASELINE.json configs[0] (bundled liver fixture, LRI_mouse, YOUNG vs OLD): counts and distributions."""
    case = helpers.liver_case(liver, mode, score_type)
    got = _check_uploaded(case["prob"], 200, seed=5, score_type=score_type, dist=True)
    assert got["stats"]["kernel_launches"] > 0


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_distributions_across_several_batches(liver_cases, mode):
    """return_distributions with more replicates than one device batch: every batch lands at its own offset."""
    prob = liver_cases[mode]["prob"]
    pcond, pct = helpers.reference_like_labels(prob, 300, 21)
    want, want_d = oracle_c.perm_counts(prob, pct, pcond, return_distributions=True)
    got = perm_counts(prob, 300, perm_cond=pcond, perm_ct=pct, batch_size=128, return_distributions=True)
    assert np.array_equal(got["counts"], want, equal_nan=True)
    assert np.array_equal(got["distributions"], want_d)
    # device-drawn labels: distributions and counts of the same call are consistent with each other
    dev = perm_counts(prob, 300, seed=8, batch_size=128, return_distributions=True)
    d, obs = dev["distributions"], np.asarray(prob.observed).reshape(prob.n_rows, prob.ncols)
    if mode == "nocond":
        assert np.array_equal((d[:, 0, :] >= obs[:, 0]).sum(0), dev["counts"][:, 0])
    else:
        assert np.array_equal((np.abs(d[:, 0, :]) >= np.abs(obs[:, 0])).sum(0), dev["counts"][:, 0])
        assert np.array_equal((d[:, 1, :] >= obs[:, 1]).sum(0), dev["counts"][:, 1])
        assert np.array_equal((d[:, 2, :] >= obs[:, 2]).sum(0), dev["counts"][:, 2])


def test_liver_counts_do_not_depend_on_batch_size(liver_cases):
    prob = liver_cases["cond"]["prob"]
    pcond, pct = helpers.reference_like_labels(prob, 300, 9)
    a = perm_counts(prob, 300, perm_cond=pcond, perm_ct=pct, batch_size=128)["counts"]
    b = perm_counts(prob, 300, perm_cond=pcond, perm_ct=pct, batch_size=384)["counts"]
    assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_group_means_bit_exact(liver_cases, mode):
    """aggregate_cells on the original labels: means and detection rates (reference R/utils_cci.R:441-462, 262-266)."""
    prob = liver_cases[mode]["prob"]
    grp = np.asarray(prob.ct_code) + (prob.n_celltypes * np.asarray(prob.cond_code) if mode == "cond" else 0)
    with Engine(prob) as eng:
        means, dr = eng.group_means(detection_rate=True)
    assert np.array_equal(means, oracle_c.group_means(prob, grp))
    assert np.array_equal(dr, oracle_c.group_means(prob, grp, detection=True))


@pytest.mark.parametrize("urn", [0, 1])
@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_device_philox_labels_match_cpu_restatement(liver_cases, mode, urn, monkeypatch):
    """urn = 0: the random-key generators (K1k / K1k2); urn = 1: the sequential urn generator (SCDC_URN_LABELS=1 selects
    it in the library and in the oracle's restatement alike)."""
    monkeypatch.setenv("SCDC_URN_LABELS", str(urn))
    prob = liver_cases[mode]["prob"]
    with Engine(prob) as eng:
        cond, ct = eng.draw_labels(200, seed=1234, perm_offset=1000)
    cond_o, ct_o = oracle_c.philox_labels(prob, 200, seed=1234, perm_offset=1000)
    assert np.array_equal(ct, ct_o)
    if mode == "cond":
        assert np.array_equal(cond, cond_o)


@pytest.mark.parametrize("kw", [
    dict(n_cells=4000, n_genes=30, n_celltypes=30, n_conditions=2, n_lri=40, seed=11, min_group=10),    # NG = 4
    dict(n_cells=8000, n_genes=30, n_celltypes=60, n_conditions=2, n_lri=40, seed=12, min_group=10),    # NG = 8
    dict(n_cells=6000, n_genes=20, n_celltypes=100, n_conditions=1, n_lri=30, seed=13, min_group=10),   # NG = 16
    dict(n_cells=9000, n_genes=20, n_celltypes=200, n_conditions=1, n_lri=30, seed=14, min_group=10),   # NG = 32
    dict(n_cells=4003, n_genes=20, n_celltypes=7, n_conditions=2, n_lri=30, seed=15, min_group=10),     # N % 4 != 0
    dict(n_cells=5001, n_genes=20, n_celltypes=9, n_conditions=1, n_lri=30, seed=16, min_group=10),     # N % 4 != 0
    dict(n_cells=70001, n_genes=30, n_celltypes=3, n_conditions=2, n_lri=40, seed=17, min_group=10),    # large strata
])
@pytest.mark.parametrize("urn", [0, 1])
def test_device_philox_labels_match_cpu_restatement_all_urn_widths(kw, urn, monkeypatch):
    monkeypatch.setenv("SCDC_URN_LABELS", str(urn))
    prob, ai, simple, mask = synthetic.make_problem(kw)
    with Engine(prob) as eng:
        cond, ct = eng.draw_labels(40, seed=kw["seed"], perm_offset=2**33 + 5)   # 64-bit replicate index
    cond_o, ct_o = oracle_c.philox_labels(prob, 40, seed=kw["seed"], perm_offset=2**33 + 5)
    assert np.array_equal(ct, ct_o)
    if cond is not None:
        assert np.array_equal(cond, cond_o)


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_device_generated_run_equals_oracle_on_the_same_labels(liver_cases, mode):
    """Philox mode end to end: counts from on-device labels == oracle counts on the CPU restatement of those labels,
    and sharding the replicate range (perm_offset) does not change the total."""
    prob = liver_cases[mode]["prob"]
    n, seed = 333, 77
    cond_o, ct_o = oracle_c.philox_labels(prob, n, seed=seed)
    want, _ = oracle_c.perm_counts(prob, ct_o, cond_o)
    with Engine(prob) as eng:
        whole = eng.run(n, seed=seed)["counts"]
        a = eng.run(200, seed=seed, perm_offset=0)["counts"]
        b = eng.run(n - 200, seed=seed, perm_offset=200)["counts"]
    assert np.array_equal(whole, want, equal_nan=True)
    assert np.array_equal(np.nan_to_num(a) + np.nan_to_num(b), np.nan_to_num(want))


@pytest.mark.parametrize("mode", ["cond", "nocond"])
def test_label_super_batches_and_pipelining_do_not_change_counts(liver_cases, mode, monkeypatch):
    """Labels are drawn per super-batch (normally one); with several, K1 runs one ahead on a second stream."""
    prob = liver_cases[mode]["prob"]
    with Engine(prob) as eng:
        whole = eng.run(1000, seed=5, batch_size=128)["counts"]
        monkeypatch.setenv("SCDC_SUPER_BATCH", "256")
        piped = eng.run(1000, seed=5, batch_size=128)["counts"]
        monkeypatch.setenv("SCDC_SUPER_BATCH", "128")
        piped2 = eng.run(1000, seed=5, batch_size=128)["counts"]
    assert np.array_equal(whole, piped, equal_nan=True) and np.array_equal(whole, piped2, equal_nan=True)


def test_uploaded_labels_that_are_not_stratified_shuffles_are_rejected(liver_cases):
    prob = liver_cases["cond"]["prob"]
    pcond, pct = helpers.reference_like_labels(prob, 4, 1)
    pcond[2, :10] = 1 - pcond[2, :10]
    with pytest.raises(_lib.ScdcError, match="do not preserve"):
        perm_counts(prob, 4, perm_cond=pcond, perm_ct=pct)
    pct[1, 0] = 200
    with pytest.raises(_lib.ScdcError, match="outside"):
        perm_counts(prob, 4, perm_cond=helpers.reference_like_labels(prob, 4, 1)[0], perm_ct=pct)


@pytest.mark.parametrize("kw", [
    dict(n_cells=3000, n_genes=200, n_celltypes=6, n_conditions=2, n_lri=300, seed=1),     # K=12  -> J=4
    dict(n_cells=4000, n_genes=150, n_celltypes=12, n_conditions=2, n_lri=250, seed=2),    # K=24  -> J=2
    dict(n_cells=9000, n_genes=120, n_celltypes=30, n_conditions=2, n_lri=200, seed=3),    # K=60  -> J=1
    dict(n_cells=2500, n_genes=180, n_celltypes=9, n_conditions=1, n_lri=300, seed=4),     # detection, K=9
    dict(n_cells=20000, n_genes=60, n_celltypes=60, n_conditions=2, n_lri=100, seed=5, min_group=20),  # K=120
])
def test_synthetic_uploaded_labels_bit_exact(kw):
    prob, ai, simple, mask = synthetic.make_problem(kw)
    _check_uploaded(prob, 160, seed=kw["seed"], dist=(kw["seed"] == 1))


def test_run_stat_analysis_mirror(liver_cases):
    """The reference's own permutation tests (tests/testthat/test-everything.R:504-578): non-p-value columns unchanged,
    p-value columns added with 1 outside the tested rows and NA for absent subunits, and p-values identical with and
    without return_distributions under the same seed."""
    case = liver_cases["cond"]
    ai, simple = case["ai"], case["simple"]
    a = run_stat_analysis(ai, simple, 10, False, "geometric_mean", False, seed=123)
    b = run_stat_analysis(ai, simple, 10, True, "geometric_mean", False, seed=123)
    for k, v in simple.items():
        assert np.array_equal(a["cci_raw"][k], v, equal_nan=True)
    for name in ("P_VALUE_DE", "P_VALUE_YOUNG", "P_VALUE_OLD", "L1_P_VALUE_DE", "L2_P_VALUE_DE", "R1_P_VALUE_DE",
                 "R2_P_VALUE_DE", "R3_P_VALUE_DE"):
        assert np.array_equal(a["cci_raw"][name], b["cci_raw"][name], equal_nan=True), name
    p = a["cci_raw"]["P_VALUE_DE"]
    assert np.all(p[~case["mask"]] == 1.0) and np.all((p >= 0) & (p <= 1))
    assert np.isnan(a["cci_raw"]["L2_P_VALUE_DE"]).sum() > 0 and not np.isnan(a["cci_raw"]["L1_P_VALUE_DE"]).any()
    d = b["distributions"]
    assert set(d) == {"DISTRIBUTIONS_DE", "DISTRIBUTIONS_YOUNG", "DISTRIBUTIONS_OLD"}
    assert d["DISTRIBUTIONS_DE"].shape == (8952, 11)
    obs = d["DISTRIBUTIONS_DE"][:, -1]
    pv = (np.abs(d["DISTRIBUTIONS_DE"][:, :10]) >= np.abs(obs)[:, None]).sum(1) / 10
    assert np.array_equal(pv, a["cci_raw"]["P_VALUE_DE"][case["mask"]])
    with pytest.raises(ValueError, match="1000"):
        run_stat_analysis(ai, simple, 1500, True, "geometric_mean", False)


def test_device_shuffles_are_uniform(liver_cases):
    """Statistical parity with base::sample: every cell is equally likely to receive every label (chi-square on the
    per-cell label frequencies over 4000 replicates)."""
    prob = liver_cases["nocond"]["prob"]
    n = 4000
    with Engine(prob) as eng:
        _, ct = eng.draw_labels(n, seed=99)
    T, N = prob.n_celltypes, prob.n_cells
    p = np.bincount(prob.ct_code, minlength=T) / N
    obs = np.stack([(ct == t).sum(0) for t in range(T)], axis=1)        # [N, T]
    chi2 = ((obs - n * p) ** 2 / (n * p)).sum()
    dof = N * (T - 1)
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)
    # adjacent cells are (nearly) uncorrelated: P(same label) = sum_t n_t (n_t - 1) / (N (N - 1))
    nt = np.bincount(prob.ct_code, minlength=T)
    expect = (nt * (nt - 1)).sum() / (N * (N - 1))
    same = (ct[:, :-1] == ct[:, 1:]).mean()
    assert abs(same - expect) < 5 * np.sqrt(expect * (1 - expect) / (n * (N - 1)))


def test_device_shuffles_are_uniform_differential(liver_cases):
    """Differential mode (random keys, two passes): cond' is a uniform shuffle inside every cell type, ct' a uniform
    shuffle inside every permuted condition; both keep n(cell type, condition)."""
    prob = liver_cases["cond"]["prob"]
    n = 3000
    with Engine(prob) as eng:
        cond, ct = eng.draw_labels(n, seed=123)
    T, N = prob.n_celltypes, prob.n_cells
    ct0, cd0 = np.asarray(prob.ct_code), np.asarray(prob.cond_code)
    base = np.zeros((2, T), np.int64)
    np.add.at(base, (cd0, ct0), 1)
    for q in (0, 1, n - 1):          # group sizes of both passes are preserved
        p1 = np.zeros((2, T), np.int64); np.add.at(p1, (cond[q], ct0), 1)
        p2 = np.zeros((2, T), np.int64); np.add.at(p2, (cond[q], ct[q]), 1)
        assert np.array_equal(p1, base) and np.array_equal(p2, base)
    # pass 1: P(cond' = 1 | cell of type t) = n(t, cond2) / n_t for every cell
    p1 = (base[1] / base.sum(0))[ct0]                       # [N]
    freq = cond.mean(0)
    z = (freq - p1) / np.sqrt(p1 * (1 - p1) / n)
    assert np.abs(z).max() < 5.5 and abs((z ** 2).sum() - N) < 6 * np.sqrt(2 * N)
    # pass 2: P(ct' = t | cond' = c) = n(t, c) / n_c, pooled over cells
    for c in (0, 1):
        sel = cond == c
        obs = np.array([(ct[sel] == t).sum() for t in range(T)])
        exp = base[c] / base[c].sum() * sel.sum()
        assert np.allclose(obs, exp)                        # exact by construction (every replicate keeps the counts)
    # ... and per cell: chi-square of the per-cell ct' frequencies against the mixture over cond'
    pc = np.stack([1 - p1, p1], axis=1)                     # [N, 2] P(cond' = c) for the cell
    pt = pc @ (base / base.sum(1, keepdims=True))           # [N, T]
    obs = np.stack([(ct == t).sum(0) for t in range(T)], axis=1)
    chi2 = ((obs - n * pt) ** 2 / (n * pt)).sum()
    dof = N * (T - 1)
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)


@pytest.mark.parametrize("no_nccl", [0, 1])
def test_in_library_multi_gpu_equals_single_gpu(liver_cases, no_nccl, monkeypatch):
    """scdc_perm_counts with n_gpus = 2: the problem is uploaded once and broadcast over NVLink (ncclBroadcast, or peer
    copies with SCDC_NO_NCCL=1), host thread per device, replicate ranges, ncclAllReduce of the counters (or peer copies +
    adds)."""
    from scdiffcom_b200 import device_count
    if device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    monkeypatch.setenv("SCDC_NO_NCCL", str(no_nccl))
    for mode in ("cond", "nocond"):
        prob = liver_cases[mode]["prob"]
        one = perm_counts(prob, 1000, seed=3, n_gpus=1)
        two = perm_counts(prob, 1000, seed=3, n_gpus=2)
        assert two["stats"]["n_gpus"] == 2 and two["stats"]["used_nccl"] == 1 - no_nccl
        assert two["stats"]["p2p_bytes"] > prob.nnz * 12 and two["stats"]["h2d_bytes"] < 1.2 * one["stats"]["h2d_bytes"]
        assert np.array_equal(one["counts"], two["counts"], equal_nan=True)
        f1 = perm_counts(prob, 1000, seed=3, n_gpus=1, precision="fp32_fast")
        f2 = perm_counts(prob, 1000, seed=3, n_gpus=2, precision="fp32_fast")
        assert np.array_equal(f1["counts"], f2["counts"], equal_nan=True)
        assert np.array_equal(f1["band_counts"], f2["band_counts"], equal_nan=True)
        pcond, pct = helpers.reference_like_labels(prob, 300, 4)
        a = perm_counts(prob, 300, perm_cond=pcond, perm_ct=pct, n_gpus=1)["counts"]
        b = perm_counts(prob, 300, perm_cond=pcond, perm_ct=pct, n_gpus=2)["counts"]
        assert np.array_equal(a, b, equal_nan=True)
    # the interrupt callback is polled by the calling thread only; the per-device worker threads watch its flag
    with pytest.raises(_lib.ScdcError, match="SCDC_EINTERRUPT"):
        perm_counts(prob, 400_000, seed=3, n_gpus=2, batch_size=128, interrupt=lambda: True)
    again = perm_counts(prob, 1000, seed=3, n_gpus=2)
    assert np.array_equal(one["counts"], again["counts"], equal_nan=True)


def test_de_column_filter_is_exact():
    """The DE column is decided from rsqrt.approx when its error band cannot change the outcome and from the IEEE sqrt
    otherwise: 50 M adversarial (x1, x0, |observed|) triples, thresholds within ulps of the true difference included."""
    import ctypes as C
    lib = _lib.load()
    out = (C.c_uint64 * 3)()
    _lib.check(lib.scdc_selftest_de_filter(50_000_000, 12345, out))
    mismatches, fast, bad_approx = int(out[0]), int(out[1]), int(out[2])
    assert mismatches == 0
    assert bad_approx == 0
    assert fast > 10_000_000          # the fast path decides whenever the threshold is not within the band


@pytest.mark.parametrize("mode", ["cond", "nocond"])
@pytest.mark.parametrize("score_type", ["geometric_mean", "arithmetic_mean"])
def test_observed_pass_on_gpu_and_one_resident_engine(liver, mode, score_type):
    """SURVEY.md "next" f2: run_simple_cci_analysis(compute_fast = FALSE) on the GPU equals the host / oracle tables bit for
    bit (expression, detection rate, CCI_SCORE*, IS_CCI_EXPRESSED*: the reference's known-answer columns,
    tests/testthat/test-everything.R:377-490), and the same resident engine then runs the permutation test on the rows it
    found (scdc_set_rows) with the counts of the one-shot call."""
    from oracle import oracle_np
    from scdiffcom_b200 import cci, preprocessing
    ai = helpers.liver_inputs(liver, mode, preprocessing)
    tmpl = cci.create_cci_template(ai)
    host = cci.run_simple_cci_analysis(ai, tmpl, score_type=score_type)
    ai_o = helpers.liver_inputs(liver, mode, oracle_np)
    want = oracle_np.run_simple_cci_analysis(ai_o, oracle_np.create_cci_template(ai_o), score_type=score_type)
    with Engine(cci.base_problem(ai)) as eng:
        dev = cci.run_simple_cci_analysis(ai, tmpl, score_type=score_type, engine=eng)
        assert set(dev) == set(host)
        # against the ORACLE's table directly (string-keyed restatement of R/utils_cci.R:171-439), same row order
        names = np.array(ai["cell_types"])
        assert (names[dev["EMITTER_CODE"]] == want["EMITTER_CELLTYPE"]).all() and (ai["LRI"]["LRI"][dev["LRI_IDX"]] == want["LRI"]).all()
        n_cmp = 0
        for k, v in dev.items():
            if k in want and np.asarray(v).dtype.kind in "fb":
                if "LOGFC" in k:
                    np.testing.assert_allclose(v, want[k], rtol=1e-15, atol=0, equal_nan=True)
                else:
                    assert np.array_equal(v, want[k], equal_nan=True), k
                n_cmp += 1
        assert n_cmp >= (9 if mode == "nocond" else 20)
        for k, v in host.items():
            if k in ("LOGFC", "LOGFC_ABS"):
                np.testing.assert_allclose(dev[k], v, rtol=1e-15, atol=0)
            else:
                assert np.array_equal(dev[k], v, equal_nan=True), k
        a = run_stat_analysis(ai, dev, 300, False, score_type, False, seed=11, engine=eng)
    b = run_stat_analysis(ai, host, 300, False, score_type, False, seed=11)
    assert int(a["tested"].sum()) == (8952 if mode == "cond" else 7352)
    for name in b["cci_raw"]:
        if "P_VALUE" in name:
            assert np.array_equal(a["cci_raw"][name], b["cci_raw"][name], equal_nan=True), name


@pytest.mark.parametrize("name", ["config2", "config3"])
def test_full_size_identity_labels_reproduce_the_observed_pass(name):
    """BASELINE.json's full-size synthetic configs through a size-independent property: uploading the ORIGINAL labels as
    every replicate must reproduce the observed scores exactly, and `>=` counts ties, so every counter equals the number of
    replicates (NaN where the subunit is NA). Observed values come from the GPU observed pass on the same engine."""
    from scdiffcom_b200 import cci
    from scdiffcom_b200 import permutation as perm
    kw = dict(synthetic.CONFIGS[name]); kw.pop("iterations")
    ai = synthetic.make_analysis_inputs(**kw)
    tmpl = cci.create_cci_template(ai)
    n = 128
    with Engine(cci.base_problem(ai)) as eng:
        simple = cci.run_simple_cci_analysis(ai, tmpl, engine=eng)
        mask, r_lri, r_emit, r_recv, obs = perm.tested_row_arrays(ai, simple)
        eng.set_rows(r_lri, r_emit, r_recv, obs)
        pr = eng.problem
        pct = np.tile(np.asarray(pr.ct_code, dtype=np.uint8), (n, 1))
        pcond = np.tile(np.asarray(pr.cond_code, dtype=np.uint8), (n, 1)) if pr.n_conditions == 2 else None
        got = eng.run(n, perm_cond=pcond, perm_ct=pct)["counts"]
        # device-drawn replicates: every counter stays within [0, n] and the run is deterministic
        a = eng.run(n, seed=3)["counts"]
        b = eng.run(n, seed=3)["counts"]
    expect = np.where(np.isnan(obs), np.nan, float(n))
    assert got.shape == expect.shape and got.shape[0] > 200_000
    assert np.array_equal(got, expect, equal_nan=True)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.nanmax(a) <= n and np.nanmin(a) >= 0
    zero_obs = obs[:, -1 if pr.n_conditions == 1 else 2] == 0          # observed score 0: every replicate counts
    assert np.all(a[zero_obs, -1 if pr.n_conditions == 1 else 2] == n)


def test_one_shot_detection_call_reduces_batch_0_behind_the_chunked_upload(liver_cases, monkeypatch):
    """scdc_perm_counts, detection mode: the matrix is copied gene chunk by gene chunk with the group sums of batch 0
    launched behind every chunk. Forced to many small chunks on the small fixture; results must not change."""
    prob = liver_cases["nocond"]["prob"]
    cond_o, ct_o = oracle_c.philox_labels(prob, 700, seed=31)
    want, _ = oracle_c.perm_counts(prob, ct_o, cond_o)
    monkeypatch.setenv("SCDC_CHUNK_NNZ", "5000")
    a = perm_counts(prob, 700, seed=31, batch_size=256)
    monkeypatch.setenv("SCDC_NO_EARLY_REDUCE", "1")
    b = perm_counts(prob, 700, seed=31, batch_size=256)
    assert np.array_equal(a["counts"], want, equal_nan=True)
    assert np.array_equal(b["counts"], want, equal_nan=True)
    n_chunks = min(16, prob.nnz // 5000)
    assert n_chunks >= 8
    assert a["stats"]["reduce_launches"] == n_chunks + 2 and b["stats"]["reduce_launches"] == 3


def test_one_shot_differential_call_starts_the_pass2_reduction_before_the_rows_are_installed(liver_cases, monkeypatch):
    """scdc_perm_counts, differential mode: batch 0's pass-2 scatter is launched right after the matrix upload and the
    pass-1 copy / rows go up on the copy stream behind it (`launch_early_scatter`). Same counts as the oracle on the
    CPU restatement of the labels, as the plain path, and as a resident engine."""
    prob = liver_cases["cond"]["prob"]
    cond_o, ct_o = oracle_c.philox_labels(prob, 700, seed=37)
    want, _ = oracle_c.perm_counts(prob, ct_o, cond_o)
    a = perm_counts(prob, 700, seed=37, batch_size=256)
    monkeypatch.setenv("SCDC_NO_EARLY_REDUCE", "1")
    b = perm_counts(prob, 700, seed=37, batch_size=256)
    monkeypatch.delenv("SCDC_NO_EARLY_REDUCE")
    with Engine(prob) as eng:
        c = eng.run(700, seed=37)
    for got in (a, b, c):
        assert np.array_equal(got["counts"], want, equal_nan=True)
    assert a["stats"]["reduce_launches"] == b["stats"]["reduce_launches"] == 2 * 3   # pass 1 + pass 2 per batch of 256


def test_interrupt_callback_stops_a_run_and_leaves_the_engine_usable(liver_cases):
    """scdc_options.interrupt (what the .Call wrapper wires to R_CheckUserInterrupt): polled on the calling thread while
    at most two batches are in flight; a truthy return ends the call with SCDC_EINTERRUPT. A callback that never fires
    changes nothing."""
    prob = liver_cases["cond"]["prob"]
    polls = []
    with Engine(prob) as eng:
        want = eng.run(4096, seed=8, batch_size=128)["counts"]
        same = eng.run(4096, seed=8, batch_size=128, interrupt=lambda: polls.append(1) and False)["counts"]
        assert np.array_equal(want, same, equal_nan=True)
        with pytest.raises(_lib.ScdcError, match="SCDC_EINTERRUPT"):
            eng.run(200_000, seed=8, batch_size=128, interrupt=lambda: True)
        again = eng.run(4096, seed=8, batch_size=128)["counts"]
        assert np.array_equal(want, again, equal_nan=True)
    with pytest.raises(_lib.ScdcError, match="SCDC_EINTERRUPT"):
        perm_counts(prob, 200_000, seed=8, batch_size=128, interrupt=lambda: True)
