"""CPU tests (no GPU): the C oracle against the committed golden vectors (outputs of the
reference's own kernels, tests/golden/make_golden.py) and, where oracle/_ref is built, against
the reference kernels executed live on fresh seeded inputs."""
import glob
import os

import numpy as np
import pytest

from mfrec_b200 import synth
from oracle import cpu, ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def g(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "kmf_*.npz"))))
def test_kmf_golden(path):
    d = dict(np.load(path))
    kernel = "linear" if "linear" in os.path.basename(path) else "logistic"
    u, v = d["u0"].copy(), d["v0"].copy()
    ib, ub = np.zeros(u.shape[1]), np.zeros(v.shape[1])
    cpu.kmf_train(kernel, int(d["nbr_epochs"]), int(d["k"]), float(d["lr"]), float(d["K_users"]),
                  float(d["K_items"]), float(d["K_bias"]), u, v, d["idx"], d["r"], ib, ub,
                  int(d["update_users"]), int(d["update_items"]))
    for a, b in ((u, d["u"]), (v, d["v"]), (ib, d["ib"]), (ub, d["ub"])):
        assert np.array_equal(a, b), "C oracle must be bit-exact with the reference kernel"


def test_funk_golden():
    d = g("funk_without_bias.npz")
    k, ni, nu = int(d["k"]), d["u"].shape[1], d["v"].shape[1]
    u = np.zeros((k, ni)) + float(d["f_init"])
    v = np.zeros((k, nu)) + float(d["f_init"])
    cpu.funk_train("without_bias", int(d["min_epochs"]), float(d["min_improvement"]), k,
                   float(d["f_init"]), float(d["lr"]), float(d["K"]), u, v, d["idx"], d["r"])
    assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])

    d = g("funk_with_bias.npz")
    u = np.zeros((k, ni)) + float(d["f_init"])
    v = np.zeros((k, nu)) + float(d["f_init"])
    cpu.funk_train("with_bias", int(d["min_epochs"]), float(d["min_improvement"]), k,
                   float(d["f_init"]), float(d["lr"]), float(d["K"]), u, v, d["idx"], d["r"],
                   float(d["mu"]), d["bi"], d["bu"])
    assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])

    for tag in ("u1_i0", "u0_i1"):
        d = g("funk_with_bias_dev_%s.npz" % tag)
        u, v = d["u0"].copy(), d["v0"].copy()
        cpu.funk_train("with_bias_dev", int(d["min_epochs"]), float(d["min_improvement"]), k,
                       float(d["f_init"]), float(d["lr"]), float(d["K"]), u, v, d["idx"], d["r"],
                       float(d["mu"]), d["bi"], d["bu"], int(d["update_users"]), int(d["update_items"]))
        assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])
        if tag == "u1_i0":
            assert np.array_equal(u, d["u0"])  # items frozen
        else:
            assert np.array_equal(v, d["v0"])


def test_predictors_and_rmse_golden():
    d = g("predictors.npz")
    for name in cpu.PREDICTORS:
        out = cpu.predict_pairs(name, d["u"], d["v"], d["pairs"], float(d["mu"]), d["ib"], d["ub"])
        np.testing.assert_allclose(out, d["pred_" + name], rtol=1e-13, atol=1e-14)
        stats, errs = cpu.rmse_pairs(name, d["u"], d["v"], d["pairs"], d["real"], float(d["mu"]),
                                     d["ib"], d["ub"])
        np.testing.assert_allclose(stats, d["stats_" + name], rtol=1e-12)
        assert stats[3] == d["pairs"].shape[0] - 1 and np.isnan(errs[17])


def test_topn_golden():
    d = g("topn.npz")
    N = int(d["N"])
    for tag, predictor, ncand in (("gd_all", "predict_rating", d["u"].shape[1]),
                                  ("mf_first25", "predict_logistic", 25)):
        for row, user in enumerate(d["users"]):
            a, b = d["rated_indptr"][row], d["rated_indptr"][row + 1]
            items, scores = cpu.topn_user(predictor, d["u"], d["v"], int(user), int(ncand),
                                          d["rated_items"][a:b], N, float(d["mu"]), d["ib"], d["ub"])
            want = d["items_" + tag][row]
            assert np.array_equal(items, want[want >= 0])
            np.testing.assert_allclose(scores, d["scores_" + tag][row][: len(items)], rtol=1e-13)
            # quirk: the item whose id equals the user id is never recommended
            assert int(user) not in items.tolist()


def test_bias_stats_against_numpy():
    nu, ni, nnz = 80, 60, 1500
    d = synth.make_ratings(nu, ni, nnz, seed=4, shuffle_seed=None)
    idx, r = d["idx"], d["r"]
    mu, ib, ub = cpu.bias_stats(idx, r, ni, nu, 0.02, 0.03)
    assert abs(mu - r.mean()) < 1e-14
    for i in range(ni):
        m = idx[:, 1] == i
        want = (r[m] - mu).sum() / (0.03 + m.sum()) if m.any() else 0.0
        assert abs(ib[i] - want) < 1e-12
    for j in range(nu):
        m = idx[:, 0] == j
        want = (r[m] - mu - ib[idx[m, 1]]).sum() / (0.02 + m.sum()) if m.any() else 0.0
        assert abs(ub[j] - want) < 1e-12


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_equals_reference_live(seed):
    """Fresh inputs, reference kernels executed here: bit-equality for every loop."""
    nu, ni, nnz, k = 120, 90, 3000, 10
    d = synth.make_ratings(nu, ni, nnz, seed=seed, shuffle_seed=seed + 7)
    idx, r = d["idx"], d["r"]
    for kernel in ("linear", "logistic"):
        for gates in ((1, 1), (0, 1), (1, 0)):
            u, v = synth.init_factors(nu, ni, k, seed=seed)
            ib, ub = np.zeros(ni), np.zeros(nu)
            u2, v2, ib2, ub2 = u.copy(), v.copy(), ib.copy(), ub.copy()
            fn = getattr(ref.kmf_train(), "train_%s_kernel" % kernel)
            fn(3, k, 0.1, 0.01, 0.0, 0.0, 0.04, 0.06, 0.007, 2.0, u, v, idx, r, ib, ub, gates[0], gates[1], 0)
            cpu.kmf_train(kernel, 3, k, 0.01, 0.04, 0.06, 0.007, u2, v2, idx, r, ib2, ub2, *gates)
            assert np.array_equal(u, u2) and np.array_equal(v, v2)
            assert np.array_equal(ib, ib2) and np.array_equal(ub, ub2)
    mu, bi, bu = cpu.bias_stats(idx, r, ni, nu)
    u = np.zeros((4, ni)) + 0.1
    v = np.zeros((4, nu)) + 0.1
    u2, v2 = u.copy(), v.copy()
    ref.gd_estimator().estimator_loop_with_bias(5, 9, 0.001, 4, 0.1, 0.002, 0, 0, 0.05, mu, u, v, idx, r,
                                                bi, bu, nu, ni, 0)
    cpu.funk_train("with_bias", 5, 0.001, 4, 0.1, 0.002, 0.05, u2, v2, idx, r, mu, bi, bu)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)


def test_synth_is_deterministic_and_unique():
    a = synth.make_ratings(200, 150, 4000, seed=0)
    b = synth.make_ratings(200, 150, 4000, seed=0)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["r"], b["r"])
    keys = a["idx"][:, 0].astype(np.int64) * 150 + a["idx"][:, 1]
    assert np.unique(keys).shape[0] == 4000
    assert a["r"].min() >= 1 and a["r"].max() <= 5


def test_als_wrmf_port_matches_reference_kernel():
    """oracle_als_wrmf (Gauss-Jordan solve) against the reference's own als_wrmf
    (als_implicit.pyx:208-352, numpy.linalg.inv) built into oracle/_ref: 1e-9."""
    from scipy.sparse import lil_matrix

    from mfrec_b200.lib.datasets import create_bool_sparse_col, create_bool_sparse_row
    from oracle import cpu, ref
    mod = ref.als_implicit()
    if mod is None:
        pytest.skip("oracle/_ref/als_implicit not built")
    rng = np.random.default_rng(4)
    nu, ni, k = 23, 17, 6
    m = lil_matrix((nu, ni))
    for _ in range(120):
        m[int(rng.integers(nu - 2)), int(rng.integers(ni - 1))] = 1.0     # trailing rows / cols stay empty
    ur, uc = create_bool_sparse_row(m)
    ir, ic = create_bool_sparse_col(m)
    # the arrays are what the reference's own helpers produce (lib/datasets.py:13-32)
    rows, cols = m.nonzero()
    assert np.array_equal(ur, np.r_[0, np.bincount(rows)]) and np.array_equal(uc, cols.astype(np.int32))
    u0 = rng.normal(0, 0.1, (k, ni))
    v0 = rng.normal(0, 0.1, (k, nu))
    u1, v1 = u0.copy(), v0.copy()
    mod.als_wrmf(3, k, u0, v0, np.zeros((k, k)), np.zeros((k, k)), ur, uc, ir, ic, nu, ni, 1, 0.015, 0)
    cpu.als_wrmf(3, k, u1, v1, ur, uc, ir, ic, 1, 0.015)
    np.testing.assert_allclose(u1, u0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(v1, v0, rtol=1e-9, atol=1e-12)
    assert not np.array_equal(v0[:, :5], np.zeros((k, 5)))


def test_funk_dev_variants_golden():
    """A3 development loops (gd_estimator.pyx:210-303, 308-395, 903-995, 401-483): the C
    restatement against vectors made by the reference's own kernels, bit for bit."""
    d = dict(np.load(os.path.join(GOLD, "funk_dev.npz")))
    idx, r = d["idx"], d["r"]
    k, f_init, lr, K = int(d["k"]), float(d["f_init"]), float(d["lr"]), float(d["K"])
    ni, nu = d["loop_u"].shape[1], d["loop_v"].shape[1]

    def fresh():
        return np.zeros((k, ni)) + f_init, np.zeros((k, nu)) + f_init

    u, v = fresh()
    hist = np.zeros_like(d["loop_hist"])
    cpu.funk_loop_dev(int(d["loop_min_epochs"]), int(d["loop_max_epochs"]), float(d["loop_min_improvement"]),
                      k, f_init, lr, K, u, v, idx, r, int(d["loop_batch"]), hist)
    assert np.array_equal(u, d["loop_u"]) and np.array_equal(v, d["loop_v"]) and np.array_equal(hist, d["loop_hist"])

    u, v = fresh()
    cpu.funk_loop_dev(int(d["loop2_min_epochs"]), -1, float(d["loop2_min_improvement"]), k, f_init, lr, K,
                      u, v, idx, r)
    assert np.array_equal(u, d["loop2_u"]) and np.array_equal(v, d["loop2_v"])

    u, v = fresh()
    cache = np.zeros(nu * ni)
    rm = []
    for f in range(2):
        for _ in range(3):
            rm.append(cpu.funk_subloop(f, k, f_init, lr, K, u, v, idx, r, cache))
        cpu.funk_predictor_subloop(f, k, f_init, u, v, idx, cache)
    assert np.array_equal(np.array(rm), d["sub_rmse"])
    assert np.array_equal(u, d["sub_u"]) and np.array_equal(v, d["sub_v"])
    assert np.array_equal(cache[d["sub_cache_cells"]], d["sub_cache_values"]) and cache.sum() == float(d["sub_cache_sum"])

    u, v = fresh()
    ib, ub = d["lb_ib0"].copy(), d["lb_ub0"].copy()
    cpu.funk_learned_bias(int(d["lb_min_epochs"]), float(d["lb_min_improvement"]), k, f_init, lr,
                          float(d["lb_lr_users"]), float(d["lb_lr_items"]), K, float(d["lb_K_bias"]),
                          float(d["lb_mu"]), u, v, idx, r, ib, ub)
    assert np.array_equal(u, d["lb_u"]) and np.array_equal(v, d["lb_v"])
    assert np.array_equal(ib, d["lb_ib"]) and np.array_equal(ub, d["lb_ub"])
