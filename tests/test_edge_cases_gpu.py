"""GPU parity on small adversarial problems built directly on the C-ABI boundary: empty gene columns, no tested rows,
2 cell types, odd sizes, absent subunits, observed values that tie EXACTLY with a replicate (the `>=` of
reference R/utils_permutation.R:141-143,164-200 counts ties), zeros and negative observed differences."""
import numpy as np
import pytest

from oracle import oracle_c
from scdiffcom_b200 import Engine, PermutationProblem, perm_counts
from tests import helpers

pytestmark = pytest.mark.gpu


random_problem = helpers.random_problem
with_tying_observed = helpers.with_tying_observed


CASES = [
    dict(seed=1, C=1, N=37, G=3, T=2, L=2, nL=1, nR=1, R=5),
    dict(seed=2, C=2, N=64, G=7, T=2, L=4, nL=2, nR=3, R=40),
    dict(seed=3, C=2, N=211, G=12, T=5, L=9, nL=2, nR=2, R=300),
    dict(seed=4, C=1, N=500, G=20, T=7, L=15, nL=1, nR=3, R=700),
    dict(seed=5, C=2, N=333, G=5, T=3, L=6, nL=1, nR=1, R=0),          # nothing to test
    dict(seed=6, C=1, N=129, G=4, T=4, L=3, nL=2, nR=1, R=1),
    dict(seed=7, C=2, N=1000, G=9, T=17, L=20, nL=2, nR=3, R=1500),    # K = 34 -> J = 1
    dict(seed=8, C=1, N=90, G=6, T=30, L=8, nL=1, nR=2, R=100, empty_cols=False),
    dict(seed=9, C=2, N=150, G=10, T=4, L=5, nL=3, nR=4, R=60),        # wider complexes than the reference's 2 + 3
]


@pytest.mark.parametrize("kw", CASES, ids=[f"case{c['seed']}" for c in CASES])
@pytest.mark.parametrize("score_type", ["geometric_mean", "arithmetic_mean"])
def test_small_adversarial_problems_bit_exact(kw, score_type):
    prob, rng = random_problem(**kw)
    n = 150                                                            # not a multiple of 32 / 128
    pcond, pct = with_tying_observed(prob, rng, n)
    st = 0 if score_type == "geometric_mean" else 1
    want, want_d = oracle_c.perm_counts(prob, pct, pcond, score_type=st, return_distributions=True)
    got = perm_counts(prob, n, score_type=score_type, perm_cond=pcond, perm_ct=pct)
    assert np.array_equal(got["counts"], want, equal_nan=True)
    if prob.n_rows:
        assert (want[:, 0] >= 1).all() or score_type == "arithmetic_mean"   # replicate 0 ties with observed -> counted
    got_d = perm_counts(prob, n, score_type=score_type, perm_cond=pcond, perm_ct=pct, return_distributions=True)
    assert np.array_equal(got_d["counts"], want, equal_nan=True)
    assert np.array_equal(got_d["distributions"], want_d)
    # device-drawn labels, any split of the replicate range
    cond_o, ct_o = oracle_c.philox_labels(prob, 77, seed=kw["seed"], perm_offset=5)
    want2, _ = oracle_c.perm_counts(prob, ct_o, cond_o, score_type=st)
    with Engine(prob) as eng:
        a = eng.run(30, seed=kw["seed"], perm_offset=5, score_type=score_type)["counts"]
        b = eng.run(47, seed=kw["seed"], perm_offset=35, score_type=score_type)["counts"]
    assert np.array_equal(np.nan_to_num(a) + np.nan_to_num(b), np.nan_to_num(want2))
