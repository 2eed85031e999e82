"""GPU parity of the Funk-SVD per-feature loops (gd_estimator.pyx:489-779) through the C ABI.
Both schedules run in float64 with unfused arithmetic, so they are compared for equality:
sequential == reference order (golden vectors made by the reference's own kernels),
stratified == oracle replay of the stratified block order."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def native():
    from mfrec_b200 import _native
    _native.default_context()
    return _native


def test_sequential_matches_reference_golden(native):
    from mfrec_b200.lib import gd_estimator
    from mfrec_b200.lib._buffers import options
    options["schedule"] = "sequential"
    try:
        d = dict(np.load(os.path.join(GOLD, "funk_without_bias.npz")))
        k, ni, nu = int(d["k"]), d["u"].shape[1], d["v"].shape[1]
        u = np.zeros((k, ni)) + float(d["f_init"])
        v = np.zeros((k, nu)) + float(d["f_init"])
        gd_estimator.estimator_loop_without_bias(int(d["min_epochs"]), 99, float(d["min_improvement"]), k,
                                                 float(d["f_init"]), float(d["lr"]), float(d["K"]), u, v,
                                                 d["idx"], d["r"], nu, ni, 0)
        assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])
        d = dict(np.load(os.path.join(GOLD, "funk_with_bias.npz")))
        u = np.zeros((k, ni)) + float(d["f_init"])
        v = np.zeros((k, nu)) + float(d["f_init"])
        gd_estimator.estimator_loop_with_bias(int(d["min_epochs"]), 99, float(d["min_improvement"]), k,
                                              float(d["f_init"]), float(d["lr"]), 0.3, 0.3, float(d["K"]),
                                              float(d["mu"]), u, v, d["idx"], d["r"], d["bi"], d["bu"], nu, ni)
        assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])
        for tag in ("u1_i0", "u0_i1"):
            d = dict(np.load(os.path.join(GOLD, "funk_with_bias_dev_%s.npz" % tag)))
            u, v = d["u0"].copy(), d["v0"].copy()
            gd_estimator.estimator_loop_with_bias_dev(
                int(d["min_epochs"]), 99, float(d["min_improvement"]), k, float(d["f_init"]), float(d["lr"]),
                0.3, 0.3, float(d["K"]), float(d["mu"]), u, v, d["idx"], d["r"], d["bi"], d["bu"], nu, ni,
                int(d["update_users"]), int(d["update_items"]), 0)
            assert np.array_equal(u, d["u"]) and np.array_equal(v, d["v"])
    finally:
        options["schedule"] = "stratified"


@pytest.mark.parametrize("variant", ["without_bias", "with_bias", "with_bias_dev"])
@pytest.mark.parametrize("B,W", [(1, 1), (2, 4), (3, 8)])
def test_stratified_equals_oracle_replay(native, small_problem, variant, B, W):
    from oracle import cpu
    p = small_problem
    k, f_init, lr, K, min_ep, min_imp = 5, 0.1, 0.002, 0.05, 4, 0.0005
    R = native.Ratings(p["idx"], p["r"], p["ni"], p["nu"], row_blocks=B, workers=W, keep_order=1, k_hint=4)
    rep = R.replay_order()
    idx_r, r_r = np.ascontiguousarray(p["idx"][rep]), np.ascontiguousarray(p["r"][rep])
    mu, bi, bu = cpu.bias_stats(p["idx"], p["r"], p["ni"], p["nu"])
    gates = (1, 0) if variant == "with_bias_dev" else (1, 1)
    u0 = np.zeros((k, p["ni"])) + f_init
    v0 = np.zeros((k, p["nu"])) + f_init
    u1, v1 = u0.copy(), v0.copy()
    _, fe_o, fr_o = cpu.funk_train(variant, min_ep, min_imp, k, f_init, lr, K, u0, v0, idx_r, r_r,
                                   mu, bi, bu, *gates)
    vid = {"without_bias": 0, "with_bias": 1, "with_bias_dev": 2}[variant]
    fe_g, fr_g = native.train_funk(vid, min_ep, 99, min_imp, k, f_init, lr, K, mu, u1, v1, p["idx"], p["r"],
                                   bi, bu, gates[0], gates[1], row_blocks=B, workers=W)
    assert np.array_equal(fe_g, fe_o), (fe_g, fe_o)
    np.testing.assert_allclose(fr_g, fr_o, rtol=1e-12)
    assert np.array_equal(u1, u0) and np.array_equal(v1, v0), "stratified Funk must equal the replay bit for bit"


def test_funk_dropin_converges_ml100k(native, ml100k_problem):
    """GDRecommender.train's kernel on configs[0]: end-of-training RMSE within 0.5 % of the
    reference order."""
    from mfrec_b200.lib import gd_estimator
    from oracle import cpu
    p = ml100k_problem
    k = 6
    u0 = np.zeros((k, p["ni"])) + 0.1
    v0 = np.zeros((k, p["nu"])) + 0.1
    u1, v1 = u0.copy(), v0.copy()
    _, fe, fr = cpu.funk_train("without_bias", 10, 0.0001, k, 0.1, 0.003, 0.05, u0, v0, p["idx"], p["r"])
    gd_estimator.estimator_loop_without_bias(10, 10, 0.0001, k, 0.1, 0.003, 0.05, u1, v1, p["idx"], p["r"],
                                             p["nu"], p["ni"], 0)
    so, _ = cpu.rmse_pairs("predict_rating", u0, v0, p["probe_idx"], p["probe_r"])
    sg, _ = cpu.rmse_pairs("predict_rating", u1, v1, p["probe_idx"], p["probe_r"])
    assert abs(gd_estimator.last_feature_rmse[-1] - fr[-1]) / fr[-1] < 5e-3
    assert abs(sg[0] - so[0]) / so[0] < 5e-3, (sg[0], so[0])


def test_dev_variants_match_reference_golden(native):
    """A3 development loops through the drop-in module (reference signatures): equal, bit for
    bit, to vectors made by the reference's own kernels."""
    from mfrec_b200.lib import gd_estimator
    d = dict(np.load(os.path.join(GOLD, "funk_dev.npz")))
    idx, r = d["idx"], d["r"]
    k, f_init, lr, K = int(d["k"]), float(d["f_init"]), float(d["lr"]), float(d["K"])
    ni, nu = d["loop_u"].shape[1], d["loop_v"].shape[1]

    def fresh():
        return np.zeros((k, ni)) + f_init, np.zeros((k, nu)) + f_init

    u, v = fresh()
    hist = np.zeros_like(d["loop_hist"])
    assert gd_estimator.estimator_loop(int(d["loop_min_epochs"]), int(d["loop_max_epochs"]),
                                       float(d["loop_min_improvement"]), k, f_init, lr, K, u, v, idx, r,
                                       int(d["loop_batch"]), hist, nu, ni, 0) is None
    assert np.array_equal(u, d["loop_u"]) and np.array_equal(v, d["loop_v"]) and np.array_equal(hist, d["loop_hist"])

    u, v = fresh()
    gd_estimator.estimator_loop2(int(d["loop2_min_epochs"]), 99, float(d["loop2_min_improvement"]), k, f_init,
                                 lr, K, u, v, idx, r, np.zeros(ni), nu, ni, 0)
    assert np.array_equal(u, d["loop2_u"]) and np.array_equal(v, d["loop2_v"])

    u, v = fresh()
    cache = np.zeros(nu * ni)
    rm = []
    for f in range(2):
        for _ in range(3):
            rm.append(gd_estimator.estimator_subloop(f, 1, 0.0, k, f_init, lr, K, u, v, idx, r, cache, nu, ni, 0))
        gd_estimator.predictor_subloop(f, 1, k, f_init, u, v, idx, r, cache, nu, ni)
    assert np.array_equal(np.array(rm), d["sub_rmse"])
    assert np.array_equal(u, d["sub_u"]) and np.array_equal(v, d["sub_v"])
    assert np.array_equal(cache[d["sub_cache_cells"]], d["sub_cache_values"]) and cache.sum() == float(d["sub_cache_sum"])

    u, v = fresh()
    ib, ub = d["lb_ib0"].copy(), d["lb_ub0"].copy()
    gd_estimator.estimator_loop_with_learned_bias(
        int(d["lb_min_epochs"]), 99, float(d["lb_min_improvement"]), k, f_init, lr, float(d["lb_lr_users"]),
        float(d["lb_lr_items"]), K, float(d["lb_K_bias"]), float(d["lb_mu"]), u, v, idx, r, ib, ub, nu, ni, 0)
    assert np.array_equal(u, d["lb_u"]) and np.array_equal(v, d["lb_v"])
    assert np.array_equal(ib, d["lb_ib"]) and np.array_equal(ub, d["lb_ub"])

    with pytest.raises(NotImplementedError):
        gd_estimator.estimator_loop_with_implicit_feedback()
    with pytest.raises(ValueError):   # the dense cache is indexed with nbr_users: it must be v's width
        gd_estimator.estimator_subloop(0, 1, 0.0, k, f_init, lr, K, u, v, idx, r, cache, nu + 1, ni, 0)
