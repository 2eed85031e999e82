"""Size-independent properties at BASELINE.json's FULL size (configs[2]: Netflix shape, 480,000 x 17,700,
100 M ratings, k = 128), where the oracle cannot follow (run with ``-m gpu``; ~1 minute).

The oracle comparisons of tests/test_sgd_gpu.py run on 6,000 ratings and those of
tests/test_convergence_gpu.py on up to 20 M; this file checks, on the whole headline problem, what
must hold at any size:

* the packer is a bijection that carries every triple bit-exactly (ratings/index preprocessing must
  be bit-exact, north_star), every bucket is sorted by (user, item), every user belongs to ONE row
  group and every (virtual) item to ONE column group, and the bucket a rating sits in is the one the
  schedule assigns to that pair of groups -- which, with the Latin-square schedule of sgd.cu (in
  sub-epoch s and phase p worker (rb, w) holds column group ((rb + s) mod B, (w + p) mod W)), is
  conflict freedom;
* training is deterministic (two runs from the same state are bit-identical), the running RMSE
  falls epoch over epoch, and the error of the trained model measured by ``predict_kernel`` on the
  resident model equals the one-shot host-array path on a sample.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def netflix():
    import torch
    import bench
    from mfrec_b200 import _native, synth
    nu, ni, nnz, k = synth.SHAPES["netflix"]
    dev = torch.device("cuda", 0)
    idx_d, r_d = bench.gpu_synth(torch, dev, nu, ni, nnz, seed=0)
    ctx = _native.default_context(0)
    R = _native.Ratings(None, None, ni, nu, ctx=ctx, device_ptrs=(idx_d.data_ptr(), r_d.data_ptr()), nnz=nnz,
                        ratings_are_f32=True, k_hint=k, keep_order=1)
    yield dict(nu=nu, ni=ni, nnz=nnz, k=k, idx_d=idx_d, r_d=r_d, R=R, ctx=ctx)
    R.close()


def test_packer_properties_at_full_size(netflix):
    p = netflix
    R, nu, ni, nnz = p["R"], p["nu"], p["ni"], p["nnz"]
    B, W, G = R.B, R.W, R.G
    idx = p["idx_d"].cpu().numpy()
    r = p["r_d"].cpu().numpy()
    up, ip = R.perms()
    assert np.array_equal(np.sort(up), np.arange(nu))
    vbase, item_rows, _n_split = R.copies()          # hot items are trained as several virtual items
    assert ip.min() >= 0 and ip.max() < item_rows and np.unique(ip).size == ni
    order = R.order()
    pu, pi, pr = R.packed()
    pu &= 0x07ffffff                                  # (the hint bits above the ids are the kernel's)
    pi &= 0x07ffffff
    valid = order >= 0
    # bijection: every input rating exactly once
    assert int(valid.sum()) == nnz
    seen = np.zeros(nnz, dtype=bool)
    seen[order[valid]] = True
    assert seen.all()
    del seen
    src = order[valid]
    # triples carried bit-exactly: user relabelled by the permutation, rating as is
    assert np.array_equal(pu[valid], up[idx[src, 0]])
    # items: a packed row belongs to ONE item (a hot item is trained as several rows, its copies), an item
    # has at most as many rows as the layout gave it copies, and ip[] names one of them
    it = idx[src, 1]
    item_of_row = np.full(item_rows, -1, dtype=np.int32)
    item_of_row[pi[valid]] = it
    assert (item_of_row[pi[valid]] == it).all(), "two items share a packed row"
    rows_used = np.bincount(item_of_row[item_of_row >= 0], minlength=ni)
    assert (rows_used <= (vbase[1:] - vbase[:-1])).all()
    rated = rows_used > 0
    assert (item_of_row[ip[rated]] == np.flatnonzero(rated)).all()
    assert np.array_equal(pr[valid], r[src])
    del it, src
    # buckets: counts, alignment, (user, item) order inside every bucket
    off, cnt = R.offsets()
    assert int(cnt.sum()) == nnz and (off[:-1] % 4 == 0).all() and int(cnt.max()) == R.max_bucket
    nb = R.n_buckets
    bucket = np.repeat(np.arange(nb, dtype=np.int64), cnt)            # bucket of the j-th valid entry (packed order)
    assert np.array_equal(np.flatnonzero(valid), (np.repeat(off[:-1], cnt) + (np.arange(nnz) - np.repeat(np.cumsum(cnt) - cnt, cnt))))
    key = pu[valid].astype(np.int64) * item_rows + pi[valid]
    same = bucket[1:] == bucket[:-1]
    assert (np.diff(key)[same] > 0).all(), "a bucket must be sorted by (user, item)"
    del key, same
    # every user in ONE row group, every virtual item in ONE column group, and the bucket of a rating is
    # the one the schedule gives that pair of groups
    q = bucket.copy()
    ph = q % W; q //= W
    w = q % W; q //= W
    cbl = q % B; q //= B
    rb = q % B
    slab = q // B
    del q
    row_group = (rb * W + w).astype(np.int32)
    col_group = ((slab * B + cbl) * W + (w + ph) % W).astype(np.int32)
    rg_of_user = np.full(nu, -1, dtype=np.int32)
    rg_of_user[pu[valid]] = row_group
    assert (rg_of_user[pu[valid]] == row_group).all(), "a user's ratings sit in more than one row group"
    cg_of_item = np.full(item_rows, -1, dtype=np.int32)
    cg_of_item[pi[valid]] = col_group
    assert (cg_of_item[pi[valid]] == col_group).all(), "an item's ratings sit in more than one column group"
    # groups are contiguous ranges of packed ids (the Q tile of a column block is one bulk copy)
    assert (np.diff(rg_of_user[rg_of_user >= 0]) >= 0).all()
    assert (np.diff(cg_of_item[cg_of_item >= 0]) >= 0).all()


def test_training_is_deterministic_and_converges_at_full_size(netflix):
    import torch
    from mfrec_b200 import _native, synth
    p = netflix
    R, nu, ni, nnz, k, ctx = p["R"], p["nu"], p["ni"], p["nnz"], p["k"], p["ctx"]
    u0, v0 = synth.init_factors(nu, ni, k, seed=2)
    hp = (0.005, 0.05, 0.05, 0.007)
    epochs = 4

    def run():
        M = _native.Model(k, ni, nu, u0, v0, None, None, layout=R, ctx=ctx)
        se = torch.zeros(epochs, device="cuda", dtype=torch.float64)
        for e in range(epochs):
            M.sgd_epoch(R, _native.KERNEL_LINEAR, *hp, sq_err_ptr=se.data_ptr() + 8 * e)
        ctx.sync()
        torch.cuda.synchronize()
        return M, torch.sqrt(se / nnz).cpu().numpy()

    Ma, ca = run()
    a = Ma.read()
    # the error of the trained, resident model on a sample of the training pairs ...
    n_s = 2_000_000
    _, st = Ma.predict("predict_linear", None, want_stats=True,
                       device_ptrs=(p["idx_d"].data_ptr(), p["r_d"].data_ptr(), 0), n=n_s, real_is_f32=True)
    del Ma
    Mb, cb = run()
    b = Mb.read()
    del Mb
    assert np.array_equal(ca, cb)
    for x, y in zip(a, b):
        assert np.array_equal(x, y), "two runs from the same state must be bit-identical"
    assert np.isfinite(ca).all() and (np.diff(ca) < 0).all(), ca
    for x in a:
        assert np.isfinite(x).all()
    # ... equals the one-shot path on the downloaded factors (float64 host arrays in)
    idx_s = p["idx_d"][:n_s].cpu().numpy()
    r_s = p["r_d"][:n_s].double().cpu().numpy()
    stats, _ = _native.rmse_pairs("predict_linear", a[0], a[1], idx_s, r_s, 0.0, a[2], a[3])
    rmse_resident = float(np.sqrt(st[0] / st[2]))
    assert abs(stats[0] - rmse_resident) <= 1e-5 * rmse_resident
    assert rmse_resident < ca[0]
