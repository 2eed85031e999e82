"""GPU parity of the evaluation entry points (predict pairs, RMSE, top-N, bias statistics) through
the C ABI, against the golden vectors and the CPU oracle.  Floating point: the device computes in
fp32, so predictions must agree within 1e-5 relative (north_star), written below."""
import os

import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


@pytest.fixture(scope="module")
def native():
    from mfrec_b200 import _native
    _native.default_context()
    return _native


def test_predictors_and_rmse_golden(native):
    d = dict(np.load(os.path.join(GOLD, "predictors.npz")))
    for name in native.PREDICTORS:
        out = native.predict_pairs(name, d["u"], d["v"], d["pairs"], float(d["mu"]), d["ib"], d["ub"])
        np.testing.assert_allclose(out, d["pred_" + name], rtol=RTOL, atol=RTOL)
        stats, errs = native.rmse_pairs(name, d["u"], d["v"], d["pairs"], d["real"], float(d["mu"]), d["ib"], d["ub"])
        np.testing.assert_allclose(stats[:2], d["stats_" + name][:2], rtol=RTOL)
        np.testing.assert_allclose(stats[2], d["stats_" + name][2], rtol=1e-3, atol=1e-6)
        assert stats[3] == d["stats_" + name][3]          # the NaN error is dropped
        assert np.isnan(errs[17]) and not np.isnan(np.delete(errs, 17)).any()


@pytest.mark.parametrize("k", [5, 32, 64, 100, 128, 256])
def test_predict_all_row_widths(native, k):
    from oracle import cpu
    nu, ni = 70, 50
    u, v = synth.init_factors(nu, ni, k, seed=k)
    rng = np.random.default_rng(k)
    ib, ub = rng.normal(0, 0.3, ni), rng.normal(0, 0.3, nu)
    pairs = np.stack([rng.integers(0, nu, 500), rng.integers(0, ni, 500)], axis=1).astype(np.int32)
    real = rng.integers(1, 6, 500).astype(np.float64)
    for name in ("predict_rating_with_bias", "predict_logistic"):
        got = native.predict_pairs(name, u, v, pairs, 3.5, ib, ub)
        want = cpu.predict_pairs(name, u, v, pairs, 3.5, ib, ub)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=RTOL)
        sg, _ = native.rmse_pairs(name, u, v, pairs, real, 3.5, ib, ub)
        so, _ = cpu.rmse_pairs(name, u, v, pairs, real, 3.5, ib, ub)
        np.testing.assert_allclose(sg[:2], so[:2], rtol=RTOL)


def test_predict_edge_cases(native):
    u, v = synth.init_factors(6, 4, 3)
    out = native.predict_pairs("predict_rating", u, v, np.zeros((0, 2), np.int32))
    assert out.shape == (0,)
    stats, errs = native.rmse_pairs("predict_rating", u, v, np.zeros((0, 2), np.int32), np.zeros(0))
    assert np.isnan(stats[0]) and stats[3] == 0
    with pytest.raises(IndexError):
        native.predict_pairs("predict_rating", u, v, np.array([[6, 0]], np.int32))
    # a resident model scores the same pairs as the one-shot call
    M = native.Model(3, 4, 6, u, v)
    pairs = np.array([[0, 0], [5, 3], [2, 1]], np.int32)
    a, _ = M.predict("predict_rating", pairs)
    b = native.predict_pairs("predict_rating", u, v, pairs)
    assert np.array_equal(a, b)


def _same_ranking(items, scores, want_items, want_scores, tol=2e-5):
    """ids must match except where neighbouring oracle scores tie within fp32 resolution."""
    assert len(items) == len(want_items)
    np.testing.assert_allclose(scores, want_scores, rtol=RTOL, atol=RTOL)
    for j, (a, b) in enumerate(zip(items, want_items)):
        if a != b:
            near = [want_items[i] for i in range(len(want_items)) if abs(want_scores[i] - want_scores[j]) <= tol * max(1.0, abs(want_scores[j]))]
            assert a in near, (j, a, b)


def test_topn_golden(native):
    d = dict(np.load(os.path.join(GOLD, "topn.npz")))
    N = int(d["N"])
    for tag, predictor, ncand in (("gd_all", "predict_rating", d["u"].shape[1]),
                                  ("mf_first25", "predict_logistic", 25)):
        items, scores, counts = native.topn(predictor, d["u"], d["v"], d["users"], ncand, d["rated_indptr"],
                                            d["rated_items"], N, float(d["mu"]), d["ib"], d["ub"])
        for row in range(len(d["users"])):
            want = d["items_" + tag][row]
            m = int((want >= 0).sum())
            assert counts[row] == m
            _same_ranking(items[row][:m], scores[row][:m], want[:m], d["scores_" + tag][row][:m])
            assert (items[row][m:] == -1).all()
            assert int(d["users"][row]) not in items[row].tolist()     # id quirk


def test_topn_against_oracle(native, small_problem):
    from oracle import cpu
    p = small_problem
    k = 40
    u, v = synth.init_factors(p["nu"], p["ni"], k, seed=9)
    rng = np.random.default_rng(1)
    ib, ub = rng.normal(0, 0.2, p["ni"]), rng.normal(0, 0.2, p["nu"])
    users = rng.permutation(p["nu"])[:97].astype(np.int32)
    order = np.lexsort((p["idx"][:, 1], p["idx"][:, 0]))
    sidx = p["idx"][order]
    starts = np.searchsorted(sidx[:, 0], users)
    ends = np.searchsorted(sidx[:, 0], users + 1)
    indptr = np.zeros(len(users) + 1, np.int64)
    indptr[1:] = np.cumsum(ends - starts)
    rated = np.concatenate([sidx[a:b, 1] for a, b in zip(starts, ends)]).astype(np.int32)
    N = 20
    for predictor, ncand in (("predict_rating", p["ni"]), ("predict_linear", 130), ("predict_logistic", p["ni"])):
        items, scores, counts = native.topn(predictor, u, v, users, ncand, indptr, rated, N, 3.3, ib, ub)
        for row, user in enumerate(users):
            wi, ws = cpu.topn_user(predictor, u, v, int(user), ncand, rated[indptr[row]:indptr[row + 1]], N, 3.3, ib, ub)
            assert counts[row] == len(wi)
            _same_ranking(items[row][:len(wi)], scores[row][:len(wi)], wi, ws)
    # N larger than the candidate set: padded with -1
    items, scores, counts = native.topn("predict_rating", u, v, users[:3], 10, indptr[:4], rated[:indptr[3]], 16, 0.0, ib, ub)
    assert (counts <= 10).all() and (items[:, 10:] == -1).all()


def test_bias_stats(native, small_problem):
    from oracle import cpu
    p = small_problem
    mu_g, ib_g, ub_g = native.bias_stats(p["idx"], p["r"], p["ni"], p["nu"], 0.02, 0.03)
    mu_o, ib_o, ub_o = cpu.bias_stats(p["idx"], p["r"], p["ni"], p["nu"], 0.02, 0.03)
    assert abs(mu_g - mu_o) < 1e-12
    np.testing.assert_allclose(ib_g, ib_o, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ub_g, ub_o, rtol=1e-10, atol=1e-12)
