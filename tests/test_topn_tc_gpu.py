"""Tensor-core top-N sweep (mfrec_topn_sweep) against the exact path (mfrec_topn) and the CPU
oracle (mf.py:144-193 / gradient_descent.py:769-802 semantics).  The sweep's tcgen05 GEMM only
selects candidates; final scores are fp32, so they must agree with the oracle within 1e-5 relative
and the ranking must be the same except between scores tied at fp32 resolution."""
import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def native():
    from mfrec_b200 import _native
    _native.default_context()
    return _native


def _rated_csr(nu, ni, per_user, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, per_user * 2, nu)
    indptr = np.zeros(nu + 1, np.int64)
    indptr[1:] = np.cumsum(lens)
    rated = np.concatenate([np.sort(rng.choice(ni, n, replace=False)) for n in lens]).astype(np.int32)
    return indptr, rated


def _check_rows(items, scores, counts, want_items, want_scores, want_counts, rows, tol=3e-5):
    for row in rows:
        m = int(want_counts[row])
        assert counts[row] == m
        np.testing.assert_allclose(scores[row][:m], want_scores[row][:m], rtol=RTOL, atol=RTOL)
        for j in range(m):
            if items[row][j] != want_items[row][j]:
                near = [want_items[row][i] for i in range(m)
                        if abs(want_scores[row][i] - want_scores[row][j]) <= tol * max(1.0, abs(want_scores[row][j]))]
                # a tie at the very end of the list may swap with an item just outside it
                assert items[row][j] in near or j >= m - 2, (row, j)
        assert (items[row][m:] == -1).all()


@pytest.mark.parametrize("k,predictor", [(128, "predict_rating"), (64, "predict_dot"), (40, "predict_linear"),
                                         (128, "predict_logistic"), (200, "predict_rating_with_bias")])
def test_sweep_equals_exact_path(native, k, predictor):
    nu, ni, N = 1500, 4000, 25
    u, v = synth.init_factors(nu, ni, k, seed=k)
    rng = np.random.default_rng(k)
    ib, ub = rng.normal(0, 0.2, ni), rng.normal(0, 0.2, nu)
    indptr, rated = _rated_csr(nu, ni, 30, seed=1)
    items, scores, counts, stats = native.topn_sweep(predictor, u, v, None, ni, indptr, rated, N, 3.4, ib, ub)
    assert stats[3] == 2.0 * nu * ni * k           # the tensor-core path ran (not forwarded)
    assert stats[0] < 0.05 * nu, "too many users fell back to the exact path: %r" % (stats,)
    users = np.arange(nu, dtype=np.int32)
    wi, ws, wc = native.topn(predictor, u, v, users, ni, indptr, rated, N, 3.4, ib, ub)
    _check_rows(items, scores, counts, wi, ws, wc, range(nu))
    for row in (0, 7, nu - 1):                       # and the quirk: item id == user id is never returned
        assert row not in items[row].tolist()


def test_sweep_against_oracle_with_user_list_and_ragged_tiles(native):
    from oracle import cpu
    nu, ni, k, N = 700, 2100, 96, 17                # 2100 items: the last 64-item tile is ragged
    u, v = synth.init_factors(nu, ni, k, seed=5)
    rng = np.random.default_rng(2)
    ib, ub = rng.normal(0, 0.3, ni), rng.normal(0, 0.3, nu)
    users = rng.permutation(nu)[:333].astype(np.int32)
    full_indptr, full_rated = _rated_csr(nu, ni, 20, seed=3)
    indptr = np.zeros(len(users) + 1, np.int64)
    parts = [full_rated[full_indptr[x]:full_indptr[x + 1]] for x in users]
    indptr[1:] = np.cumsum([len(q) for q in parts])
    rated = np.concatenate(parts).astype(np.int32)
    ncand = 2050
    items, scores, counts, stats = native.topn_sweep("predict_linear", u, v, users, ncand, indptr, rated, N, 0.0, ib, ub)
    assert stats[3] > 0
    for row in range(0, len(users), 7):
        wi, ws = cpu.topn_user("predict_linear", u, v, int(users[row]), ncand, parts[row], N, 0.0, ib, ub)
        assert counts[row] == len(wi)
        np.testing.assert_allclose(scores[row][:len(wi)], ws, rtol=RTOL, atol=RTOL)
        assert (items[row][:len(wi)] < ncand).all()
        assert set(items[row][:len(wi) - 2]) <= set(wi)


def test_sweep_forwards_small_problems(native, small_problem):
    p = small_problem
    u, v = synth.init_factors(p["nu"], p["ni"], 16, seed=1)
    items, scores, counts, stats = native.topn_sweep("predict_rating", u, v, None, p["ni"], None, None, 50)
    assert stats[3] == 0                            # 16 * N > items: exact path
    wi, ws, wc = native.topn("predict_rating", u, v, np.arange(p["nu"], dtype=np.int32), p["ni"], None, None, 50)
    assert np.array_equal(items, wi) and np.array_equal(scores, ws) and np.array_equal(counts, wc)


def test_sweep_heavy_tailed_scores_still_exact(native):
    """Scores far from normal (a few huge item rows): thresholds misjudge, the certificate sends
    those users to the exact path, results stay right."""
    nu, ni, k, N = 600, 3000, 32, 30
    u, v = synth.init_factors(nu, ni, k, seed=8)
    u[:, :40] *= 25.0
    items, scores, counts, stats = native.topn_sweep("predict_dot", u, v, None, ni, None, None, N)
    wi, ws, wc = native.topn("predict_dot", u, v, np.arange(nu, dtype=np.int32), ni, None, None, N)
    _check_rows(items, scores, counts, wi, ws, wc, range(nu))
