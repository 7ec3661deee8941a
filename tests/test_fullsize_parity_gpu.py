"""Full-size GPU parity (run with `pytest -m gpu` on a B200): BASELINE.json's synthetic configs 2-5 at their real sizes,
SHUFFLED labels, counts bit-equal to the CPU oracle (`oracle_c.perm_counts`) — reference R/utils_permutation.R:139-211
(counting), :501-603 (shuffles), R/utils_cci.R:441-462 (group means), :613-653 (complex-min; config 4 uses the real
LRI_human table). The oracle runs 64 replicates per config (seconds to a minute on the box's host cores); the production
batch sizes are tied to those 64 through the additivity of the counters over replicate ranges.
"""
import numpy as np
import pytest

import bench
from oracle import oracle_c
from scdiffcom_b200 import Engine, perm_counts

pytestmark = pytest.mark.gpu
N_ORACLE = 64


@pytest.fixture(scope="module")
def workloads():
    cache = {}

    def get(name):
        if name not in cache:
            cache.clear()                      # one full-size problem in host memory at a time
            cache[name] = bench.build_workload(name)[0]
        return cache[name]
    return get


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", ["config2", "config3", "config4", "config5"])
def test_full_size_shuffled_labels_equal_the_oracle(workloads, name):
    prob = workloads(name)
    seed = 1000 + int(name[-1])
    # the oracle's restatement of the device label generator (random keys, both passes at full N), then its counts
    cond_o, ct_o = oracle_c.philox_labels(prob, N_ORACLE, seed=seed)
    want, _ = oracle_c.perm_counts(prob, ct_o, cond_o)
    assert want.shape[0] == prob.n_rows > 300_000
    assert 0 < np.nanmean(want) < N_ORACLE          # not the trivial all / nothing answer
    iters = bench.BENCH_ITERS[name]
    with Engine(prob) as eng:
        drawn = eng.run(N_ORACLE, seed=seed)                                     # labels drawn on the device
        assert _eq(drawn["counts"], want), "device-drawn labels"
        up = eng.run(N_ORACLE, perm_cond=cond_o, perm_ct=ct_o)                   # the same labels uploaded (validation mode)
        assert _eq(up["counts"], want), "uploaded labels"
        # production batch (what bench.py times): one call over `iters` replicates == [0, 64) + [64, iters)
        whole = eng.run(iters, seed=seed)
        rest = eng.run(iters - N_ORACLE, seed=seed, perm_offset=N_ORACLE)
        assert whole["stats"]["batch_size"] >= min(iters, 1024)
        assert _eq(np.nan_to_num(whole["counts"]), np.nan_to_num(want) + np.nan_to_num(rest["counts"])), "additivity"
    one = perm_counts(prob, N_ORACLE, seed=seed)                                 # one-shot C-ABI call (what .Call binds)
    assert _eq(one["counts"], want), "one-shot call"
    # SCDC_FP32_FAST on the same replicates: every difference from the fp64 oracle lies inside the reported band
    f32 = perm_counts(prob, N_ORACLE, seed=seed, precision="fp32_fast")
    diff = np.abs(np.nan_to_num(f32["counts"]) - np.nan_to_num(want))
    band = np.nan_to_num(f32["band_counts"])
    assert np.all(diff <= band), f"fp32-fast: {int((diff > band).sum())} counters differ by more than their band"
    assert band.sum() <= 0.02 * N_ORACLE * np.isfinite(want).sum()


def test_config2_production_batch_against_the_oracle_directly(workloads):
    """Config 2 is small enough for the oracle to run the whole production batch: 10 000 uploaded replicates (numpy
    shuffles standing in for base::sample) through the B = 10 112 single-batch path, every counter compared."""
    from tests import helpers
    prob = workloads("config2")
    n = bench.BENCH_ITERS["config2"]
    _, pct = helpers.reference_like_labels(prob, n, seed=7)
    want, _ = oracle_c.perm_counts(prob, pct, None)
    got = perm_counts(prob, n, perm_ct=pct)
    assert got["stats"]["batch_size"] >= 1024
    assert _eq(got["counts"], want)


@pytest.mark.parametrize("name", ["config3", "config4"])
def test_full_size_numpy_shuffled_uploaded_labels(workloads, name):
    """Uploaded labels drawn like run_stat_iteration does (numpy Generator standing in for base::sample), differential
    configs, both score types on 32 replicates."""
    from tests import helpers
    prob = workloads(name)
    pcond, pct = helpers.reference_like_labels(prob, 32, seed=21)
    for st_name, st in (("geometric_mean", 0), ("arithmetic_mean", 1)):
        want, _ = oracle_c.perm_counts(prob, pct, pcond, score_type=st)
        got = perm_counts(prob, 32, perm_cond=pcond, perm_ct=pct, score_type=st_name)
        assert _eq(got["counts"], want), st_name
