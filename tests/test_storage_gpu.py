"""fp16 / bf16 user-factor storage with float32 arithmetic (``mfrec_opts.storage``; run with ``-m gpu``).

north_star: "optionally bf16/fp16 factor storage with fp32 accumulation".  The reference computes
in float64 (kmf_train.pyx:113-118); the stated tolerances of the narrower formats against the
REFERENCE's end-of-training RMSE (tests/golden/convergence.json, made by oracle/_ref) are

    float32 rows   0.5 %   (tests/test_convergence_gpu.py)
    fp16 rows      1 %     round-to-nearest on store
    bf16 rows      2 %     stochastic rounding on store (an SGD step is below half a bf16 ulp)

on C1 (ML-100K shape, k = 20 -> skipped: 16-bit rows need k > 32) and C2 (ML-20M shape, k = 64) /
C3-prefix (Netflix shape, k = 128).  Also: predictions on a 16-bit model equal predictions on the
same factors rounded on the host (bit-level check of the widening loads), training is
deterministic, and the unsupported combinations fail loudly instead of silently training float32.
"""
import functools
import importlib.util
import json
import os

import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = {"f16": 0.01, "bf16": 0.02}

_spec = importlib.util.spec_from_file_location("make_convergence", os.path.join(HERE, "golden", "make_convergence.py"))
make_convergence = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_convergence)

with open(os.path.join(HERE, "golden", "convergence.json")) as _f:
    GOLDEN = json.load(_f)


@functools.lru_cache(maxsize=1)
def _problem(name):
    return make_convergence.problem(name)


def _round_rows(v, storage):
    """float64 [k, n] -> the values a 16-bit row holds (round-to-nearest-even), as float64."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    dt = torch.float16 if storage == "f16" else torch.bfloat16
    return t.to(dt).to(torch.float32).numpy().astype(np.float64)


@pytest.mark.parametrize("storage", ["f16", "bf16"])
@pytest.mark.parametrize("kernel", ["linear", "logistic"])
@pytest.mark.parametrize("name", ["c2", "c3p"])
def test_end_of_training_rmse_within_stated_tolerance(name, kernel, storage):
    from mfrec_b200 import _native
    from mfrec_b200.lib import kmf_train
    from mfrec_b200.lib._buffers import options
    gold = GOLDEN["%s_%s" % (name, kernel)]
    p = _problem(name)
    u, v = synth.init_factors(p["nu"], p["ni"], p["k"], seed=2)
    ib, ub = np.zeros(p["ni"]), np.zeros(p["nu"])
    fn = {"linear": kmf_train.train_linear_kernel, "logistic": kmf_train.train_logistic_kernel}[kernel]
    old = options["storage"]
    options["storage"] = storage
    try:
        fn(gold["epochs"], p["k"], 0.1, gold["lr"], 0.0, 0.0, gold["K_users"], gold["K_items"], gold["K_bias"],
           0.0, u, v, p["idx"], p["r"], ib, ub)
    finally:
        options["storage"] = old
    # what came back is exactly representable in the storage format (the rows were 16-bit in HBM)
    assert np.array_equal(v, _round_rows(v, storage))
    pred = "predict_" + kernel
    train, _ = _native.rmse_pairs(pred, u, v, p["idx"], p["r"], 0.0, ib, ub)
    probe, _ = _native.rmse_pairs(pred, u, v, p["probe_idx"], p["probe_r"], 0.0, ib, ub)
    rel_t = abs(train[0] - gold["train_rmse"]) / gold["train_rmse"]
    rel_p = abs(probe[0] - gold["probe_rmse"]) / gold["probe_rmse"]
    print("%s %s %s: train %.6f (reference %.6f, rel %.2e)  probe %.6f (reference %.6f, rel %.2e)"
          % (name, kernel, storage, train[0], gold["train_rmse"], rel_t, probe[0], gold["probe_rmse"], rel_p))
    assert rel_t <= TOL[storage] and rel_p <= TOL[storage]


@pytest.mark.parametrize("storage", ["f16", "bf16"])
@pytest.mark.parametrize("k", [40, 64, 128, 200])
def test_resident_16bit_model_round_trip_and_predict(k, storage):
    """Upload narrows with round-to-nearest, read widens exactly; predict_kernel on the 16-bit rows
    equals predict_kernel on a float32 model holding the rounded values (same fma order)."""
    from mfrec_b200 import _native
    nu, ni, nnz = 700, 300, 20000
    d = synth.make_ratings(nu, ni, nnz, seed=3)
    idx, r = d["idx"], d["r"]
    u, v = synth.init_factors(nu, ni, k, seed=4)
    ib, ub = np.random.RandomState(5).normal(0, 0.1, ni), np.random.RandomState(6).normal(0, 0.1, nu)
    R = _native.Ratings(idx, r, ni, nu, k_hint=k, storage=_native.STORAGE_F16 if storage == "f16" else _native.STORAGE_BF16)
    M = _native.Model(k, ni, nu, u, v, ib, ub, layout=R)
    u1, v1, ib1, ub1 = M.read()
    vr = _round_rows(v, storage)
    assert np.array_equal(v1, vr)
    np.testing.assert_array_equal(u1, u.astype(np.float32).astype(np.float64))
    got, _ = M.predict("predict_linear", idx)
    M32 = _native.Model(k, ni, nu, u, vr, ib, ub)
    want, _ = M32.predict("predict_linear", idx)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("storage", ["f16", "bf16"])
def test_training_on_16bit_rows_is_deterministic_and_tracks_float32(storage):
    from mfrec_b200 import _native
    nu, ni, nnz, k = 3000, 800, 200000, 64
    d = synth.make_ratings(nu, ni, nnz, seed=11)
    idx, r = d["idx"], d["r"]
    hp = (0.01, 0.05, 0.05, 0.007)
    code = _native.STORAGE_F16 if storage == "f16" else _native.STORAGE_BF16

    def run(st):
        u, v = synth.init_factors(nu, ni, k, seed=12)
        ib, ub = np.zeros(ni), np.zeros(nu)
        rm = _native.train_kmf(_native.KERNEL_LINEAR, 8, k, *hp, u, v, idx, r, ib, ub, storage=st)
        return u, v, ib, ub, rm

    a = run(code)
    b = run(code)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    f = run(_native.STORAGE_F32)
    # the running RMSE of every epoch stays on the float32 curve
    np.testing.assert_allclose(a[4], f[4], rtol=TOL[storage])
    assert a[4][-1] < a[4][0]


@pytest.mark.parametrize("storage", ["f16", "bf16"])
def test_16bit_rows_with_local_slabs(storage):
    """A catalogue too large for one tile per SM is cut into local slabs inside one persistent launch
    (the Yahoo!Music shape on one GPU): the 16-bit build walks them like the float32 one."""
    from mfrec_b200 import _native
    nu, ni, nnz, k = 2000, 1500, 150000, 64
    d = synth.make_ratings(nu, ni, nnz, seed=21)
    idx, r = d["idx"], d["r"]
    hp = (0.01, 0.05, 0.05, 0.007)
    code = _native.STORAGE_F16 if storage == "f16" else _native.STORAGE_BF16

    def run(st):
        R = _native.Ratings(idx, r, ni, nu, k_hint=k, row_blocks=3, workers=4, n_slabs=2, storage=st)
        assert R.G == 2
        u, v = synth.init_factors(nu, ni, k, seed=22)
        M = _native.Model(k, ni, nu, u, v, np.zeros(ni), np.zeros(nu), layout=R)
        for _ in range(6):
            M.sgd_epoch(R, _native.KERNEL_LINEAR, *hp)
        M.ctx.sync()
        u1, v1, ib1, ub1 = M.read()
        return _native.rmse_pairs("predict_linear", u1, v1, idx, r, 0.0, ib1, ub1)[0][0]

    got, want = run(code), run(_native.STORAGE_F32)
    assert abs(got - want) <= TOL[storage] * want, (got, want)


def test_unsupported_combinations_fail_loudly():
    from mfrec_b200 import _native
    nu, ni, nnz = 400, 300, 12000
    d = synth.make_ratings(nu, ni, nnz, seed=1)
    idx, r = d["idx"], d["r"]
    hp = (0.01, 0.05, 0.05, 0.007)
    # k <= 32: one element per lane, no 16-bit build
    u, v = synth.init_factors(nu, ni, 20, seed=2)
    with pytest.raises(_native.MfrecError):
        _native.train_kmf(_native.KERNEL_LINEAR, 1, 20, *hp, u, v, idx, r, np.zeros(ni), np.zeros(nu),
                          storage=_native.STORAGE_F16)
    # one side frozen (fold-in): float32 rows only
    u, v = synth.init_factors(nu, ni, 64, seed=2)
    R = _native.Ratings(idx, r, ni, nu, k_hint=64, storage=_native.STORAGE_BF16)
    M = _native.Model(64, ni, nu, u, v, np.zeros(ni), np.zeros(nu), layout=R)
    with pytest.raises(_native.MfrecError):
        M.sgd_epoch(R, _native.KERNEL_LINEAR, *hp, update_users=0, update_items=1)
    # an unknown storage code
    with pytest.raises(ValueError):
        _native.Ratings(idx, r, ni, nu, k_hint=64, storage=7)
    # the DSGD ring trains float32 rows
    deg = np.bincount(idx[:, 1], minlength=ni).astype(np.int64)
    Rw = _native.Ratings(idx, r, ni, nu, row_blocks=2, workers=4, n_slabs=2, k_hint=64, item_degree=deg,
                         storage=_native.STORAGE_F16)
    Mw = _native.Model(64, ni, nu, u, v, None, None, layout=Rw)
    with pytest.raises(_native.MfrecError):
        _native.PeerRing(Rw, Mw, 0, 2)
