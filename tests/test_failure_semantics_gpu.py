"""Failure semantics of the GPU path (run with ``-m gpu``).

The reference computes in float64 and simply lets a diverging run overflow: its per-epoch RMSE
(kmf_train.pyx:273) becomes inf / NaN.  The stratified kernel reduces the dot product in 32-bit
fixed point, where a NaN partial converts to 0 -- so non-finite factors are detected where every
updated item row passes anyway (the tile write-back) and the epoch's RMSE is reported as NaN.
Also here: run-to-run determinism of the bias statistics (a segmented reduction, no floating
point atomics).
"""
import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu


def _problem():
    d = synth.make_ratings(300, 200, 6000, seed=0, shuffle_seed=3)
    u, v = synth.init_factors(300, 200, 16, seed=2)
    return d["idx"], d["r"], u, v, np.zeros(200), np.zeros(300)


def test_divergence_reports_nan_like_the_reference():
    from mfrec_b200 import _native
    from oracle import cpu
    idx, r, u, v, ib, ub = _problem()
    uo, vo, ibo, ubo = u.copy(), v.copy(), ib.copy(), ub.copy()
    lr = 50.0                                               # absurd learning rate: blows up within an epoch
    with np.errstate(all="ignore"):
        rm_ref = cpu.kmf_train("linear", 4, 16, lr, 0.05, 0.05, 0.007, uo, vo, idx, r, ibo, ubo)
    assert not np.isfinite(rm_ref[-1])                      # the reference ends in inf / NaN
    rm = _native.train_kmf(_native.KERNEL_LINEAR, 4, 16, lr, 0.05, 0.05, 0.007, u, v, idx, r, ib, ub)
    assert np.isnan(rm[-1]), rm                            # ... and so does the stratified schedule
    assert not np.isfinite(u).all()


def test_nan_in_the_inputs_surfaces():
    from mfrec_b200 import _native
    idx, r, u, v, ib, ub = _problem()
    r = r.copy()
    r[123] = np.nan                                         # a NaN rating poisons its rows, then the sums
    rm = _native.train_kmf(_native.KERNEL_LINEAR, 2, 16, 0.01, 0.05, 0.05, 0.007, u, v, idx, r, ib, ub)
    assert np.isnan(rm).all()
    idx, r, u, v, ib, ub = _problem()
    u[3, idx[0, 1]] = np.inf                                # a non-finite initial item factor that is trained on
    rm = _native.train_kmf(_native.KERNEL_LOGISTIC, 2, 16, 0.01, 0.05, 0.05, 0.007, u, v, idx, r, ib, ub)
    assert np.isnan(rm[-1])


def test_sane_training_is_unaffected_by_the_check():
    from mfrec_b200 import _native
    idx, r, u, v, ib, ub = _problem()
    rm = _native.train_kmf(_native.KERNEL_LINEAR, 5, 16, 0.01, 0.05, 0.05, 0.007, u, v, idx, r, ib, ub)
    assert np.isfinite(rm).all() and rm[-1] < rm[0]


def test_bias_statistics_are_deterministic_and_match_the_oracle():
    from mfrec_b200 import _native
    from oracle import cpu
    d = synth.make_ratings(5000, 800, 300000, seed=7, shuffle_seed=8)
    runs = [_native.bias_stats(d["idx"], d["r"], 800, 5000, 0.02, 0.03) for _ in range(3)]
    for mu, ib, ub in runs[1:]:
        assert mu == runs[0][0] and np.array_equal(ib, runs[0][1]) and np.array_equal(ub, runs[0][2])
    mu_o, ib_o, ub_o = cpu.bias_stats(d["idx"], d["r"], 800, 5000, 0.02, 0.03)
    assert abs(runs[0][0] - mu_o) <= 1e-12 * abs(mu_o)
    np.testing.assert_allclose(runs[0][1], ib_o, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(runs[0][2], ub_o, rtol=1e-10, atol=1e-12)
    with pytest.raises(IndexError):
        bad = d["idx"].copy()
        bad[5, 1] = 800
        _native.bias_stats(bad, d["r"], 800, 5000)
