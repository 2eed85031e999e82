"""The DSGD ring (``mfrec_ring_*``: persistent launches, column blocks pushed to the next rank's
copy of Q, counters at system scope) against the CPU oracle -- run with ``-m gpu``.

One GPU is enough: ``mfrec_ring_epochs_one_device`` runs all ranks of a ring as CTA groups of ONE
cooperative launch with exactly the kernel, counters and hand-over code a real ring uses (ranks
on one device must never be separate launches: nothing would guarantee they run together).
Ranks own disjoint users and, in any step, disjoint item slabs, so the ring is equivalent to the
sequential order  epoch -> step -> rank -> that rank's stratified order of slab (rank + step) mod
world, which the oracle replays (sequential semantics: kmf_train.pyx:241-273).

``test_ring_over_nvlink`` is the same check across real GPUs (one process per GPU over
torch.distributed / cudaIpc); it is skipped on a box with a single device.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, KU, KI, KB = 0.01, 0.05, 0.06, 0.007


def split_users(idx, nu, world):
    """Contiguous user slices with ~equal rating counts: bounds[w] .. bounds[w + 1]."""
    deg = np.bincount(idx[:, 0], minlength=nu)
    cum = np.concatenate([[0], np.cumsum(deg)])
    bounds = [int(np.searchsorted(cum, cum[-1] * w / world)) for w in range(world)] + [nu]
    bounds[0] = 0
    return bounds


def build_ranks(native, idx, r, nu, ni, k, world, B, W, u0, v0, **pack_opts):
    """Per rank: (row mask, first user, Ratings, Model, PeerRing), rings connected in-process."""
    bounds = split_users(idx, nu, world)
    deg_i = np.bincount(idx[:, 1], minlength=ni).astype(np.int64)
    ranks = []
    for w in range(world):
        a, b = bounds[w], bounds[w + 1]
        mine = (idx[:, 0] >= a) & (idx[:, 0] < b)
        idx_w = np.ascontiguousarray(idx[mine])
        idx_w[:, 0] -= a
        R = native.Ratings(idx_w, np.ascontiguousarray(r[mine]), ni, b - a, row_blocks=B, workers=W,
                           n_slabs=world, keep_order=1, k_hint=k, item_degree=deg_i, **pack_opts)
        M = native.Model(k, ni, b - a, u0, np.ascontiguousarray(v0[:, a:b]), None, None, layout=R)
        ranks.append(dict(mine=np.nonzero(mine)[0], first=a, n=b - a, R=R, M=M))
    for w in range(world):
        ranks[w]["ring"] = native.PeerRing(ranks[w]["R"], ranks[w]["M"], w, world)
    for w in range(world):
        ranks[w]["ring"].connect_local(ranks[(w - 1) % world]["ring"])
    return ranks


def ring_replay(ranks, idx, r, world, epochs):
    """Global input indices in the ring's equivalent sequential order."""
    per = [[rk["mine"][o] for o in rk["R"].replay_order(by_slab=True)] for rk in ranks]
    out = []
    for _ in range(epochs):
        for step in range(world):
            for w in range(world):
                out.append(per[w][(w + step) % world])
    return np.concatenate(out)


@pytest.mark.parametrize("kernel", ["linear", "logistic"])
@pytest.mark.parametrize("world,B,W,k", [(2, 2, 2, 16), (3, 2, 4, 40), (4, 3, 2, 128), (1, 3, 4, 24)])
def test_ring_on_one_device_matches_oracle_replay(world, B, W, k, kernel, small_problem):
    from mfrec_b200 import _native as native
    from oracle import cpu
    import torch
    p = small_problem
    idx, r, nu, ni = p["idx"], p["r"], p["nu"], p["ni"]
    u0, v0 = synth.init_factors(nu, ni, k, seed=2)
    ranks = build_ranks(native, idx, r, nu, ni, k, world, B, W, u0, v0)
    # every rank computed the same item partition from the global degrees
    ip0 = ranks[0]["R"].perms()[1]
    for rk in ranks[1:]:
        assert np.array_equal(rk["R"].perms()[1], ip0)
        assert (rk["R"].B, rk["R"].W) == (ranks[0]["R"].B, ranks[0]["R"].W)
    kid = {"linear": native.KERNEL_LINEAR, "logistic": native.KERNEL_LOGISTIC}[kernel]
    rings = [rk["ring"] for rk in ranks]
    se = torch.zeros(3, device="cuda", dtype=torch.float64)
    # two epochs in one launch, then a third in its own launch (the counters carry over)
    native.ring_epochs_one_device(rings, kid, LR, KU, KI, KB, n_epochs=2, sq_err_ptr=se.data_ptr())
    native.ring_epochs_one_device(rings, kid, LR, KU, KI, KB, n_epochs=1, sq_err_ptr=se.data_ptr() + 16)
    for g in rings:
        g.sync_model()
    u1 = np.zeros_like(u0)
    v1 = np.zeros_like(v0)
    ib1, ub1 = np.zeros(ni), np.zeros(nu)
    for w, rk in enumerate(ranks):
        uw, vw, ibw, ubw = rk["M"].read()
        a, b = rk["R"].slab_items(w)                      # after whole epochs rank w holds slab w
        in_slab = (ip0 >= a) & (ip0 < b)
        u1[:, in_slab], ib1[in_slab] = uw[:, in_slab], ibw[in_slab]
        v1[:, rk["first"]:rk["first"] + rk["n"]] = vw
        ub1[rk["first"]:rk["first"] + rk["n"]] = ubw
    rep = ring_replay(ranks, idx, r, world, 3)
    ib0, ub0 = np.zeros(ni), np.zeros(nu)
    rm = cpu.kmf_train(kernel, 1, k, LR, KU, KI, KB, u0, v0, np.ascontiguousarray(idx[rep]),
                       np.ascontiguousarray(r[rep]), ib0, ub0)
    for a_, b_ in ((u0, u1), (v0, v1), (ib0, ib1), (ub0, ub1)):
        np.testing.assert_allclose(b_, a_, rtol=2e-4, atol=2e-5)
    # the three epochs' error sums add up to the replay's (one pass over 3 x nnz ratings)
    np.testing.assert_allclose(float(se.sum().item()), float(rm[0]) ** 2 * rep.shape[0], rtol=1e-4)


def test_ring_with_hot_item_copies_matches_the_oracle():
    """Hot-item copies around a ring: the copies of an item live in ONE slab and travel with it; the
    rank that holds the slab at the end of an epoch averages them.  Oracle: the same block order
    over the same copies with the same per-epoch merge."""
    from mfrec_b200 import _native as native
    from oracle import cpu
    import torch
    world, B, W, k = 2, 2, 2, 32
    nu, ni, nnz = 1200, 12, 9000
    d = synth.make_ratings(nu, ni, nnz, seed=4, shuffle_seed=5)
    idx, r = d["idx"], d["r"]
    u0, v0 = synth.init_factors(nu, ni, k, seed=2)
    ranks = build_ranks(native, idx, r, nu, ni, k, world, B, W, u0, v0, split=native.SPLIT_ON, split_min_copy=16)
    vbase, rows, n_hot = ranks[0]["R"].copies()
    assert n_hot > 0
    for rk in ranks[1:]:
        assert np.array_equal(rk["R"].copies()[0], vbase)
    J = np.diff(vbase)
    item_of = np.repeat(np.arange(ni), J)
    rings = [rk["ring"] for rk in ranks]
    se = torch.zeros(2, device="cuda", dtype=torch.float64)
    native.ring_epochs_one_device(rings, native.KERNEL_LINEAR, LR, KU, KI, KB, n_epochs=2, sq_err_ptr=se.data_ptr())
    for g in rings:
        g.sync_model()
    # oracle
    vidx = idx.copy()
    for rk in ranks:
        m = rk["mine"]
        vidx[m, 1] = vbase[idx[m, 1]] + native.copy_of_user(idx[m, 0] - rk["first"], J[idx[m, 1]])   # rank-local user ids
    per = [[rk["mine"][o] for o in rk["R"].replay_order(by_slab=True)] for rk in ranks]
    u, ib = u0.copy(), np.zeros(ni)
    ub = np.zeros(nu)
    for _ in range(2):
        uv, ibv = np.ascontiguousarray(u[:, item_of]), ib[item_of].copy()
        for step in range(world):
            for w in range(world):
                o = per[w][(w + step) % world]
                if o.shape[0]:
                    cpu.kmf_train("linear", 1, k, LR, KU, KI, KB, uv, v0, np.ascontiguousarray(vidx[o]),
                                  np.ascontiguousarray(r[o]), ibv, ub)
        for i in range(ni):
            u[:, i] = uv[:, vbase[i]:vbase[i + 1]].mean(axis=1)
            ib[i] = ibv[vbase[i]:vbase[i + 1]].mean()
    ip0 = ranks[0]["R"].perms()[1]
    for w, rk in enumerate(ranks):
        uw, vw, ibw, ubw = rk["M"].read()
        a, b = rk["R"].slab_items(w)
        held = (ip0 >= a) & (ip0 < b)
        assert held.any()
        np.testing.assert_allclose(uw[:, held], u[:, held], rtol=3e-4, atol=3e-5)
        np.testing.assert_allclose(ibw[held], ib[held], rtol=3e-4, atol=3e-5)
        np.testing.assert_allclose(vw, v0[:, rk["first"]:rk["first"] + rk["n"]], rtol=3e-4, atol=3e-5)


def test_ring_reports_a_missing_neighbour(small_problem, monkeypatch):
    """A rank whose neighbour never shows up gives up after MFREC_RING_TIMEOUT_MS instead of
    hanging the GPU: NaN error sums, mfrec_ring_wait raises."""
    from mfrec_b200 import _native as native
    import torch
    monkeypatch.setenv("MFREC_RING_TIMEOUT_MS", "200")
    p = small_problem
    k = 16
    u0, v0 = synth.init_factors(p["nu"], p["ni"], k, seed=2)
    ranks = build_ranks(native, p["idx"], p["r"], p["nu"], p["ni"], k, 2, 2, 2, u0, v0)
    se = torch.zeros(1, device="cuda", dtype=torch.float64)
    ranks[0]["ring"].epochs(native.KERNEL_LINEAR, LR, KU, KI, KB, 1, se.data_ptr())   # rank 1 is never launched
    with pytest.raises(native.MfrecError):
        ranks[0]["ring"].wait()
    assert np.isnan(se.item())


@pytest.mark.parametrize("world", [2, 4])
def test_ring_over_nvlink(world, tmp_path):
    """T7 on hardware: `world` processes x `world` GPUs train ONE shared data set split by users
    through the peer-memory ring; final RMSE within 0.5 % of the single-GPU run of the same data,
    and two runs are bit-identical."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "ring.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world + os.getpid() % 100),
           os.path.join(ROOT, "tools", "ring_check.py"), "--out", out]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    import json
    with open(out) as f:
        res = json.load(f)
    assert res["bit_identical_runs"]
    assert res["ranks_agree_on_layout"]
    assert abs(res["rmse_ring"] - res["rmse_single"]) / res["rmse_single"] <= 0.005
    assert abs(res["probe_ring"] - res["probe_single"]) / res["probe_single"] <= 0.005


def test_kmf_recommender_trains_on_all_gpus_of_the_process():
    """`mfrec.lib.kmf_train.options["devices"] = [0, 1, ...]`: KMFRecommender.train (kmf.py:197-220)
    runs as a DSGD ring over the GPUs of THIS process (mfrec_train_kmf_multi, peer memory
    addressed directly).  Same data on one GPU: end-of-training RMSE within 0.5 %; two runs
    bit-identical."""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2 GPUs")
    from mfrec_b200 import _native as native
    from mfrec.lib import kmf_train
    nu, ni, nnz, k, epochs = 30000, 2500, 2_000_000, 64, 8
    d = synth.make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3, probe_frac=0.1)
    idx, r = d["idx"], d["r"]
    hp = (0.005, 0.05, 0.05, 0.007)

    def run(devices):
        u, v = synth.init_factors(nu, ni, k, seed=2)
        ib, ub = np.zeros(ni), np.zeros(nu)
        kmf_train.options["devices"] = devices
        try:
            kmf_train.train_linear_kernel(epochs, k, 0.1, hp[0], 0.0, 0.0, hp[1], hp[2], hp[3], 0.0, u, v, idx, r, ib, ub)
        finally:
            kmf_train.options["devices"] = []
        probe, _ = native.rmse_pairs("predict_linear", u, v, d["probe_idx"], d["probe_r"], 0.0, ib, ub)
        return u, v, ib, ub, float(kmf_train.last_rmse[-1]), float(probe[0])

    single = run([])
    for world in sorted({2, min(n_dev, 4)}):
        a = run(list(range(world)))
        b = run(list(range(world)))
        assert all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4]))
        assert abs(a[4] - single[4]) / single[4] <= 0.005 and abs(a[5] - single[5]) / single[5] <= 0.005, (a[4:], single[4:])
        assert np.isfinite(a[0]).all() and np.isfinite(a[1]).all()
