"""GPU parity of the KMF SGD path (train_linear_kernel / train_logistic_kernel) against the CPU
oracle, called through the C ABI (ctypes) -- run with ``pytest -m gpu`` on a B200.

T3  sequential schedule == reference order: bit-exact (linear), 1e-12 (logistic: exp)
T5  ratings layout: relabelling is a bijection, buckets hold the right triples, bit-exact
T4' stratified schedule == one-thread replay of the same schedule with the oracle
T4  end-of-training RMSE within 0.5 % of the reference order (north_star tolerance)
"""
import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu

LR, KU, KI, KB = 0.01, 0.05, 0.06, 0.007


@pytest.fixture(scope="module")
def native():
    from mfrec_b200 import _native
    _native.default_context()
    return _native


def _fresh(nu, ni, k, seed=2):
    u, v = synth.init_factors(nu, ni, k, seed)
    return u, v, np.zeros(ni), np.zeros(nu)


@pytest.mark.parametrize("kernel", ["linear", "logistic"])
def test_sequential_schedule_is_reference_order(native, small_problem, kernel):
    from oracle import cpu
    p = small_problem
    k = 12
    u0, v0, ib0, ub0 = _fresh(p["nu"], p["ni"], k)
    u1, v1, ib1, ub1 = u0.copy(), v0.copy(), ib0.copy(), ub0.copy()
    rm_o = cpu.kmf_train(kernel, 3, k, LR, KU, KI, KB, u0, v0, p["idx"], p["r"], ib0, ub0)
    kid = {"linear": native.KERNEL_LINEAR, "logistic": native.KERNEL_LOGISTIC}[kernel]
    rm_g = native.train_kmf(kid, 3, k, LR, KU, KI, KB, u1, v1, p["idx"], p["r"], ib1, ub1,
                            schedule=native.SCHED_SEQUENTIAL)
    if kernel == "linear":
        assert np.array_equal(u0, u1) and np.array_equal(v0, v1)
        assert np.array_equal(ib0, ib1) and np.array_equal(ub0, ub1)
        assert np.array_equal(rm_o, rm_g)
    else:  # libm exp vs CUDA exp differ in the last ulp
        for a, b in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1), (rm_o, rm_g)):
            np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("B,W,G", [(1, 1, 1), (3, 4, 1), (2, 8, 1), (2, 2, 3)])
def test_ratings_layout_is_bit_exact(native, small_problem, B, W, G):
    p = small_problem
    R = native.Ratings(p["idx"], p["r"], p["ni"], p["nu"], row_blocks=B, workers=W, n_slabs=G,
                       keep_order=1, k_hint=16)
    assert (R.B, R.W, R.G) == (B, W, G)
    up, ip = R.perms()
    assert np.array_equal(np.sort(up), np.arange(p["nu"]))
    assert np.array_equal(np.sort(ip), np.arange(p["ni"]))
    order = R.order()
    pu, pi, pr = R.packed()
    valid = order >= 0
    # every input rating appears exactly once
    assert np.array_equal(np.sort(order[valid]), np.arange(p["nnz"]))
    # triples are carried bit-exactly (ratings 1..5 are exact in float32)
    assert np.array_equal(pu[valid], up[p["idx"][order[valid], 0]])
    assert np.array_equal(pi[valid], ip[p["idx"][order[valid], 1]])
    assert np.array_equal(pr[valid].astype(np.float64), p["r"][order[valid]])
    assert not pu[~valid].any() and not pi[~valid].any()
    off, cnt = R.offsets()
    assert cnt.sum() == p["nnz"] and (off[:-1] % 4 == 0).all()
    assert cnt.max() == R.max_bucket
    # bucket membership: recompute each rating's bucket on the host from the group boundaries
    nb = R.n_buckets
    bucket_of_pos = np.repeat(np.arange(nb), cnt)
    pos = np.concatenate([np.arange(off[b], off[b] + cnt[b]) for b in range(nb) if cnt[b]])
    q = bucket_of_pos
    ph = q % W; q //= W
    w = q % W; q //= W
    cbl = q % B; q //= B
    rb = q % B; slab = q // B
    # users of a (row block, worker) and items of a (slab, column block, group) are id ranges
    for b in range(nb):
        if cnt[b] == 0:
            continue
        sl = slice(off[b], off[b] + cnt[b])
        keys = pu[sl].astype(np.int64) * p["ni"] + pi[sl]
        assert (np.diff(keys) > 0).all(), "bucket must be sorted by (user, item)"
    # conflict freedom: inside one (slab, sub-epoch, phase) no two buckets share a user or item
    sub_epoch = (cbl - rb) % B
    stage = ((slab * B + sub_epoch) * W + ph)
    worker_id = rb * W + w
    for st in np.unique(stage):
        m = stage == st
        for ids in (pu[pos[m]], pi[pos[m]]):
            owner = {}
            for i, wk in zip(ids.tolist(), worker_id[m].tolist()):
                assert owner.setdefault(i, wk) == wk, "two workers touch the same row in one phase"
    # replay order covers everything once
    rep = R.replay_order()
    assert np.array_equal(np.sort(rep), np.arange(p["nnz"]))


@pytest.mark.parametrize("kernel", ["linear", "logistic"])
@pytest.mark.parametrize("k,B,W", [(12, 3, 4), (40, 2, 8), (128, 4, 2), (200, 2, 4), (20, 1, 1),
                                   (64, 3, 5), (32, 2, 6), (16, 0, 0)])   # odd warp counts; 0 = the packer's own choice
def test_stratified_matches_oracle_replay(native, small_problem, kernel, k, B, W):
    """The parallel schedule is equivalent to SOME sequential order; replaying exactly that
    order with the float64 oracle must give the same factors up to fp32 round-off."""
    from oracle import cpu
    p = small_problem
    R = native.Ratings(p["idx"], p["r"], p["ni"], p["nu"], row_blocks=B, workers=W, keep_order=1,
                       k_hint=k)
    rep = R.replay_order()
    idx_r, r_r = np.ascontiguousarray(p["idx"][rep]), np.ascontiguousarray(p["r"][rep])
    u0, v0, ib0, ub0 = _fresh(p["nu"], p["ni"], k)
    M = native.Model(k, p["ni"], p["nu"], u0, v0, ib0, ub0, layout=R)
    kid = {"linear": native.KERNEL_LINEAR, "logistic": native.KERNEL_LOGISTIC}[kernel]
    epochs = 3
    for _ in range(epochs):
        M.sgd_epoch(R, kid, LR, KU, KI, KB)
    M.ctx.sync()
    u1, v1, ib1, ub1 = M.read()
    cpu.kmf_train(kernel, epochs, k, LR, KU, KI, KB, u0, v0, idx_r, r_r, ib0, ub0)
    for a, b in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1)):
        np.testing.assert_allclose(b, a, rtol=2e-4, atol=2e-5)


def test_update_gates(native, small_problem):
    """update_users / update_items gates (kmf_train.pyx:168-171, 266-269)."""
    from oracle import cpu
    p = small_problem
    k = 16
    for kernel, kid in (("linear", native.KERNEL_LINEAR), ("logistic", native.KERNEL_LOGISTIC)):
        for uu, ui in ((1, 0), (0, 1), (0, 0)):
            R = native.Ratings(p["idx"], p["r"], p["ni"], p["nu"], row_blocks=2, workers=4,
                               keep_order=1, k_hint=k)
            rep = R.replay_order()
            u0, v0, ib0, ub0 = _fresh(p["nu"], p["ni"], k)
            M = native.Model(k, p["ni"], p["nu"], u0, v0, ib0, ub0, layout=R)
            M.sgd_epoch(R, kid, LR, KU, KI, KB, update_users=uu, update_items=ui)
            M.ctx.sync()
            u1, v1, ib1, ub1 = M.read()
            cpu.kmf_train(kernel, 1, k, LR, KU, KI, KB, u0, v0, np.ascontiguousarray(p["idx"][rep]),
                          np.ascontiguousarray(p["r"][rep]), ib0, ub0, uu, ui)
            for a, b in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1)):
                np.testing.assert_allclose(b, a, rtol=2e-4, atol=2e-5)


def test_dropin_converged_rmse_ml100k(native, ml100k_problem):
    """BASELINE.json configs[0]: 943 x 1682, 100k ratings, k = 20.  Parallel SGD reorders the
    updates, so factors differ; the end-of-training RMSE (train and probe) must agree with the
    reference order within 0.5 % relative (north_star)."""
    from mfrec_b200.lib import kmf_train
    from oracle import cpu
    p = ml100k_problem
    k, epochs = p["k"], 30
    args = (0.01, 0.05, 0.05, 0.007)
    u0, v0, ib0, ub0 = _fresh(p["nu"], p["ni"], k)
    u1, v1, ib1, ub1 = u0.copy(), v0.copy(), ib0.copy(), ub0.copy()
    rm_o = cpu.kmf_train("linear", epochs, k, *args, u0, v0, p["idx"], p["r"], ib0, ub0)
    kmf_train.train_linear_kernel(epochs, k, 0.1, args[0], 0.0, 0.0, args[1], args[2], args[3], 0.0,
                                  u1, v1, p["idx"], p["r"], ib1, ub1)
    rm_g = kmf_train.last_rmse
    assert abs(rm_g[-1] - rm_o[-1]) / rm_o[-1] < 5e-3, (rm_g[-1], rm_o[-1])
    so, _ = cpu.rmse_pairs("predict_linear", u0, v0, p["probe_idx"], p["probe_r"], 0.0, ib0, ub0)
    sg, _ = cpu.rmse_pairs("predict_linear", u1, v1, p["probe_idx"], p["probe_r"], 0.0, ib1, ub1)
    assert abs(sg[0] - so[0]) / so[0] < 5e-3, (sg[0], so[0])
    # and the curve is a descent
    assert rm_g[-1] < rm_g[0]


def test_dropin_argument_errors(native, small_problem):
    """Same exception types as the reference's Cython buffer validation."""
    from mfrec_b200.lib import kmf_train
    p = small_problem
    k = 8
    u, v, ib, ub = _fresh(p["nu"], p["ni"], k)
    base = [1, k, 0.1, 0.01, 0.01, 0.01, 0.05, 0.05, 0.007, 0.0, u, v, p["idx"], p["r"], ib, ub]

    def call(pos, val):
        a = list(base)
        a[pos] = val
        return kmf_train.train_linear_kernel(*a)

    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'float64_t' but got 'float'"):
        call(10, u.astype(np.float32))
    with pytest.raises(ValueError, match="ndarray is not C-contiguous"):
        call(10, np.asfortranarray(u))
    with pytest.raises(ValueError, match=r"wrong number of dimensions \(expected 2, got 1\)"):
        call(10, u[0].copy())
    ro = u.copy()
    ro.flags.writeable = False
    with pytest.raises(ValueError, match="read-only"):
        call(10, ro)
    with pytest.raises(TypeError):
        kmf_train.train_linear_kernel(1, k)
    with pytest.raises(TypeError):
        call(10, None)
    bad = p["idx"].copy()
    bad[5, 1] = p["ni"]
    with pytest.raises(IndexError):
        call(12, bad)
    # keyword use of the trailing optionals works
    kmf_train.train_linear_kernel(*base, update_users=1, update_items=0, verbose=0)


def test_empty_and_tiny_inputs(native):
    from mfrec_b200.lib import kmf_train
    from oracle import cpu
    k, nu, ni = 4, 3, 5
    u, v = synth.init_factors(nu, ni, k)
    ib, ub = np.zeros(ni), np.zeros(nu)
    u_ref = u.copy()
    kmf_train.train_linear_kernel(2, k, 0.1, 0.01, 0, 0, 0.05, 0.05, 0.007, 0.0, u, v,
                                  np.zeros((0, 2), np.int32), np.zeros(0), ib, ub)
    assert np.array_equal(u, u_ref)
    # one rating, one user: everything serialises onto one warp
    idx = np.array([[2, 4]], np.int32)
    r = np.array([4.0])
    u2, v2, ib2, ub2 = u.copy(), v.copy(), ib.copy(), ub.copy()
    cpu.kmf_train("linear", 5, k, 0.01, 0.05, 0.05, 0.007, u2, v2, idx, r, ib2, ub2)
    kmf_train.train_linear_kernel(5, k, 0.1, 0.01, 0, 0, 0.05, 0.05, 0.007, 0.0, u, v, idx, r, ib, ub)
    np.testing.assert_allclose(u, u2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v, v2, rtol=1e-5, atol=1e-6)
    # fold-in shape: all ratings of one user, items frozen (KMFRecommender.retrain_user)
    rng = np.random.default_rng(0)
    items = rng.permutation(ni)[:4].astype(np.int32)
    idx = np.stack([np.full(4, 1, np.int32), items], axis=1)
    r = rng.integers(1, 6, 4).astype(np.float64)
    u3, v3, ib3, ub3 = u.copy(), v.copy(), ib.copy(), ub.copy()
    kmf_train.train_logistic_kernel(7, k, 0.1, 0.01, 0, 0, 0.05, 0.05, 0.007, 0.0, u, v, idx, r, ib, ub, 1, 0)
    assert np.array_equal(u, u3), "items must stay frozen when update_items = 0"
    assert not np.array_equal(v, v3)


def test_item_slabs_on_one_device(native, small_problem):
    """G = 3 item slabs processed back to back on one GPU (what one rank of a 3-GPU ring does over
    an epoch, minus the exchange) must equal the oracle replay of the same block order."""
    from oracle import cpu
    p = small_problem
    k = 24
    R = native.Ratings(p["idx"], p["r"], p["ni"], p["nu"], row_blocks=2, workers=4, n_slabs=3,
                       keep_order=1, k_hint=k)
    assert R.G == 3 and R.launches_per_epoch == 6
    bounds = [R.slab_items(c) for c in range(3)]
    assert bounds[0][0] == 0 and bounds[-1][1] == p["ni"]
    assert all(bounds[c][1] == bounds[c + 1][0] for c in range(2))
    rep = R.replay_order()
    u0, v0, ib0, ub0 = _fresh(p["nu"], p["ni"], k)
    M = native.Model(k, p["ni"], p["nu"], u0, v0, ib0, ub0, layout=R)
    for _ in range(2):
        for c in range(3):           # slab by slab, like the ring driver
            M.sgd_epoch(R, native.KERNEL_LINEAR, LR, KU, KI, KB, slab=c)
    M.ctx.sync()
    u1, v1, ib1, ub1 = M.read()
    cpu.kmf_train("linear", 2, k, LR, KU, KI, KB, u0, v0, np.ascontiguousarray(p["idx"][rep]),
                  np.ascontiguousarray(p["r"][rep]), ib0, ub0)
    for a, b in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1)):
        np.testing.assert_allclose(b, a, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("k,nu,ni,nnz,B,W", [(128, 500, 8, 3000, 1, 8),      # one item per column group: chain quads
                                              (64, 900, 16, 6000, 2, 4),      # two items per group
                                              (20, 3000, 400, 9000, 1, 2),    # sparse users: independent / clean quads
                                              (256, 700, 40, 5000, 1, 4)])
def test_quad_fast_paths_match_oracle_replay(native, k, nu, ni, nnz, B, W):
    """Layouts that make the packer emit mostly one kind of quad (chain / independent / clean), so
    every straight-line path of the kernel is compared with the oracle replay, for every row width."""
    from oracle import cpu
    d = synth.make_ratings(nu, ni, nnz, seed=k)
    R = native.Ratings(d["idx"], d["r"], ni, nu, row_blocks=B, workers=W, keep_order=1, k_hint=k)
    qt = R.quad_types()
    assert sum(qt.values()) > 0
    rep = R.replay_order()
    u0, v0, ib0, ub0 = _fresh(nu, ni, k)
    M = native.Model(k, ni, nu, u0, v0, ib0, ub0, layout=R)
    for kernel, kid in (("linear", native.KERNEL_LINEAR), ("logistic", native.KERNEL_LOGISTIC)):
        M.sgd_epoch(R, kid, LR, KU, KI, KB)
        cpu.kmf_train(kernel, 1, k, LR, KU, KI, KB, u0, v0, np.ascontiguousarray(d["idx"][rep]),
                      np.ascontiguousarray(d["r"][rep]), ib0, ub0)
    M.ctx.sync()
    u1, v1, ib1, ub1 = M.read()
    for a, b in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1)):
        np.testing.assert_allclose(b, a, rtol=3e-4, atol=3e-5)
    if ni == 8:
        assert qt["chain"] > qt["clean"] + qt["independent"], qt
    if ni == 400:
        assert qt["independent"] + qt["clean"] > qt["chain"], qt


def test_staged_pageable_copies_change_nothing(native, small_problem, monkeypatch):
    """Large pageable host arrays go through pinned bounce buffers filled by host threads
    (runtime.cu); with the size threshold lowered the same call must give the same bits."""
    from mfrec_b200.lib import kmf_train
    p = small_problem
    k = 24

    def run():
        u, v, ib, ub = _fresh(p["nu"], p["ni"], k)
        kmf_train.train_linear_kernel(2, k, 0.1, LR, 0.0, 0.0, KU, KI, KB, 0.0, u, v, p["idx"], p["r"], ib, ub)
        return u, v, ib, ub

    ref = run()
    monkeypatch.setenv("MFREC_STAGE_MIN_BYTES", "4096")
    got = run()
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def _train_with_copies(cpu, kernel, epochs, k, idx, r, vbase, u0, v0, ib0, ub0):
    """The oracle under the hot-item-copy semantics: every item trains as its copies (a rating goes
    to copy hash(user) mod copies), the copies are averaged after each epoch."""
    from mfrec_b200 import _native
    ni = u0.shape[1]
    J = np.diff(vbase)
    item_of = np.repeat(np.arange(ni), J)
    vidx = idx.copy()
    vidx[:, 1] = vbase[idx[:, 1]] + _native.copy_of_user(idx[:, 0], J[idx[:, 1]])
    u, ib = u0.copy(), ib0.copy()
    for _ in range(epochs):
        uv, ibv = np.ascontiguousarray(u[:, item_of]), ib[item_of].copy()
        cpu.kmf_train(kernel, 1, k, LR, KU, KI, KB, uv, v0, vidx, r, ibv, ub0)
        for i in range(ni):
            u[:, i] = uv[:, vbase[i]:vbase[i + 1]].astype(np.float32).mean(axis=1, dtype=np.float32) if J[i] > 1 else uv[:, vbase[i]]
            ib[i] = ibv[vbase[i]:vbase[i + 1]].mean()
    return u, v0, ib, ub0


@pytest.mark.parametrize("kernel", ["linear", "logistic"])
@pytest.mark.parametrize("k,nu,ni,nnz,B,W,G", [(128, 500, 8, 3000, 1, 8, 1), (32, 1500, 10, 12000, 2, 4, 1),
                                                (64, 1200, 12, 9000, 2, 2, 3)])
def test_hot_item_copies_match_the_oracle_with_the_same_semantics(native, kernel, k, nu, ni, nnz, B, W, G):
    """Hot-item splitting (pack.cu): items heavier than half a column group train as
    several copies merged after every epoch.  The GPU result must equal the oracle replaying the
    same block order over the same copies with the same per-epoch averaging."""
    from oracle import cpu
    d = synth.make_ratings(nu, ni, nnz, seed=4, shuffle_seed=5)
    idx, r = d["idx"], d["r"]
    R = native.Ratings(idx, r, ni, nu, row_blocks=B, workers=W, n_slabs=G, keep_order=1, k_hint=k,
                       split=native.SPLIT_ON, split_min_copy=16)
    vbase, rows, n_hot = R.copies()
    assert n_hot > 0 and rows == vbase[-1] > ni             # something was split
    R0 = native.Ratings(idx, r, ni, nu, row_blocks=B, workers=W, n_slabs=G, k_hint=k, split=native.SPLIT_OFF)
    assert R0.copies()[2] == 0
    rep = R.replay_order()
    assert np.array_equal(np.sort(rep), np.arange(nnz))
    u0, v0, ib0, ub0 = _fresh(nu, ni, k)
    M = native.Model(k, ni, nu, u0, v0, ib0, ub0, layout=R)
    kid = {"linear": native.KERNEL_LINEAR, "logistic": native.KERNEL_LOGISTIC}[kernel]
    for _ in range(3):
        M.sgd_epoch(R, kid, LR, KU, KI, KB)
    M.ctx.sync()
    u1, v1, ib1, ub1 = M.read()
    ur, vr, ibr, ubr = _train_with_copies(cpu, kernel, 3, k, np.ascontiguousarray(idx[rep]),
                                          np.ascontiguousarray(r[rep]), vbase, u0, v0, ib0, ub0)
    for a, b in ((ur, u1), (vr, v1), (ibr, ib1), (ubr, ub1)):
        np.testing.assert_allclose(b, a, rtol=3e-4, atol=3e-5)
    # predictions on the model read its merged rows (first copy) through the layout
    out, _ = M.predict("predict_linear", idx[:500])
    want = np.einsum("kn,kn->n", ur[:, idx[:500, 1]], vr[:, idx[:500, 0]]) + ibr[idx[:500, 1]] + ubr[idx[:500, 0]]
    np.testing.assert_allclose(out, want, rtol=1e-3, atol=1e-3)
