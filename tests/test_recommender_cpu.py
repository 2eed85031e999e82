"""Host-side logic of the Python 3 recommender classes (no GPU): rating store, COO extraction
order and RNG consumption (bit-exact with the reference's per-rating loop, base.py:1115-1131),
parameter plumbing and its error behaviour."""
import numpy as np
import pytest
from scipy.sparse import find


def _reference_get_ratings(matrix, randomize_order):
    """base.py:1115-1131 restated literally (per-rating Python loop)."""
    nbr_ratings = find(matrix)[2].shape[0]
    ratings = np.zeros(nbr_ratings, dtype=np.float64)
    ratings_index = np.zeros([nbr_ratings, 2], dtype=np.int32)
    cx = matrix.tocoo()
    for i, (user_index, feature_index, rating) in enumerate(zip(cx.row, cx.col, cx.data)):
        ratings_index[i] = [int(user_index), int(feature_index)]
        ratings[i] = rating
    index = np.arange(nbr_ratings)
    if randomize_order:
        np.random.shuffle(index)
    return ratings_index[index], ratings[index]


def _filled(cls, nu=30, ni=20, n=200, seed=0):
    rng = np.random.default_rng(seed)
    rec = cls(nu, ni)
    for _ in range(n):
        rec.set_item_by_id(int(rng.integers(nu)), int(rng.integers(ni)), float(rng.integers(1, 6)))
    return rec


def test_get_ratings_is_bit_exact_with_reference_loop():
    from mfrec_b200.recommendation import KMFRecommender
    rec = _filled(KMFRecommender)
    for randomize in (False, True):
        np.random.seed(11)
        want_idx, want_r = _reference_get_ratings(rec.relationship_matrix, randomize)
        after_ref = np.random.get_state()[1].copy()
        np.random.seed(11)
        got_idx, got_r = rec.get_ratings(randomize_order=randomize)
        after = np.random.get_state()[1].copy()
        assert got_idx.dtype == np.int32 and got_r.dtype == np.float64
        assert np.array_equal(got_idx, want_idx) and np.array_equal(got_r, want_r)
        assert np.array_equal(after, after_ref)      # same RNG consumption
    # users ascending, items ascending inside a user (scipy lil -> coo order)
    idx, _ = rec.get_ratings()
    key = idx[:, 0].astype(np.int64) * 1000 + idx[:, 1]
    assert (np.diff(key) > 0).all()


def test_bulk_ingestion_equals_per_rating_calls():
    from mfrec_b200.recommendation import GDRecommender
    rng = np.random.default_rng(3)
    nu, ni, n = 25, 15, 300
    idx = np.stack([rng.integers(0, nu, n), rng.integers(0, ni, n)], axis=1).astype(np.int32)
    r = rng.integers(1, 6, n).astype(np.float64)
    a, b = GDRecommender(nu, ni), GDRecommender(nu, ni)
    for (u, i), val in zip(idx, r):
        a.set_item_by_id(int(u), int(i), val)        # duplicates overwrite
    b.set_ratings(idx, r)
    ia, ra = a.get_ratings()
    ib, rb = b.get_ratings()
    assert np.array_equal(ia, ib) and np.array_equal(ra, rb)
    a.compute_overall_avg()
    assert a.overall_bias == ra.mean()


def test_parameters_and_labels():
    from mfrec_b200.recommendation import Error, GDRecommender, KMFRecommender
    rec = GDRecommender(4, 6, {'nbr_features': 7, 'regularization_model': 0.02, 'min_epochs': 3})
    assert (rec.dimensionality, rec.K, rec.min_epochs, rec.max_epochs) == (7, 0.02, 3, 275)
    with pytest.raises(Error):
        rec.set_parameters({'no_such_parameter': 1})
    k = KMFRecommender(4, 6, {'regularization_users': 0.5, 'nbr_epochs': 9})
    # reference quirk (kmf.py:39-41 vs :219): the key lands in K, training reads K_users
    assert k.K == 0.5 and k.K_users == 0.1 and k.nbr_epochs == 9
    assert k.nbr_users == 4 and k.nbr_items == 6
    assert k.users_index['user3'] == 3 and k.items_label[5] == 'item5'
    k.set_item_label(5, 'matrix')
    assert k.items_index['matrix'] == 5 and 'item5' not in k.items_index
    k.set_item_by_label('user1', 'matrix', 4)
    assert k.relationship_matrix[1, 5] == 4.0


def test_funk_dev_shims_validate_before_touching_the_device():
    """The A3 drop-ins (gd_estimator.pyx:210-303, 308-395, 401-483, 903-995) raise the reference's
    exception types for bad buffers, and ValueError where the reference would index out of
    bounds, before any device work."""
    import numpy as np
    import pytest
    from mfrec_b200.lib import gd_estimator as gd
    k, ni, nu = 2, 5, 4
    u, v = np.zeros((k, ni)) + 0.1, np.zeros((k, nu)) + 0.1
    idx = np.array([[0, 1], [3, 4]], dtype=np.int32)
    r = np.array([3.0, 5.0])
    hist = np.zeros(1 * 3 * k)
    with pytest.raises(ValueError):     # dtype mismatch, like Cython's buffer check
        gd.estimator_loop(1, 3, 0.0, k, 0.1, 0.01, 0.02, u.astype(np.float32), v, idx, r, 0, hist, nu, ni)
    with pytest.raises(ValueError):     # rmse_hist too short for batch 1
        gd.estimator_loop(1, 3, 0.0, k, 0.1, 0.01, 0.02, u, v, idx, r, 1, hist, nu, ni)
    with pytest.raises(ValueError):     # the dense cache is indexed with nbr_users: must be v's width
        gd.estimator_loop2(1, 3, 0.0, k, 0.1, 0.01, 0.02, u, v, idx, r, np.zeros(ni), nu + 1, ni)
    with pytest.raises(ValueError):     # dense cache smaller than nbr_users * nbr_items
        gd.estimator_subloop(0, 1, 0.0, k, 0.1, 0.01, 0.02, u, v, idx, r, np.zeros(3), nu, ni)
    with pytest.raises(IndexError):
        gd.predictor_subloop(k, 1, k, 0.1, u, v, idx, r, np.zeros(nu * ni), nu, ni)
    with pytest.raises(ValueError):     # read-only bias array where the loop writes it
        ib = np.zeros(ni)
        ib.setflags(write=False)
        gd.estimator_loop_with_learned_bias(1, 9, 0.0, k, 0.1, 0.01, 0.01, 0.01, 0.02, 0.01, 3.0, u, v, idx, r,
                                            ib, np.zeros(nu), nu, ni)
    with pytest.raises(NotImplementedError):
        gd.estimator_loop_with_implicit_feedback()
    assert np.all(u == 0.1) and np.all(v == 0.1)


def test_gd_development_trainers_drive_the_native_loops_like_the_reference(monkeypatch):
    """feature_training2 / feature_training_dev / feature_training_bias (gradient_descent.py:299-329,
    577-599, 472-503): host control flow and argument mapping, with the CPU oracle standing in for
    the device library (no GPU here)."""
    from mfrec_b200.lib import gd_estimator
    from mfrec_b200.recommendation import GDRecommender
    from oracle import cpu
    rng = np.random.default_rng(11)
    nu, ni, n = 30, 20, 260
    keys = rng.choice(nu * ni, n, replace=False)
    idx = np.stack([keys // ni, keys % ni], axis=1).astype(np.int32)
    r = rng.integers(1, 6, n).astype(np.float64)
    params = {'nbr_features': 3, 'min_epochs': 2, 'max_epochs': 5, 'min_improvement': 0.001,
              'learning_rate': 0.01, 'regularization_model': 0.02}

    def make():
        rec = GDRecommender(nu, ni, dict(params))
        rec.set_ratings(idx, r)
        return rec

    # feature_training2: Python epoch loop over estimator_subloop + predictor_subloop == estimator_loop2
    monkeypatch.setattr(gd_estimator, "estimator_subloop",
                        lambda f, ep, mi, dim, fi, lr, K, u, v, ri, ra, cache, nbu, nbi, verbose=0:
                        cpu.funk_subloop(f, dim, fi, lr, K, u, v, ri, ra, cache))
    monkeypatch.setattr(gd_estimator, "predictor_subloop",
                        lambda f, ep, dim, fi, u, v, ri, ra, cache, nbu, nbi:
                        cpu.funk_predictor_subloop(f, dim, fi, u, v, ri, cache))
    rec = make()
    rec.feature_training2()
    ri, ra = rec.get_ratings()
    u = np.zeros((3, ni)) + 0.1
    v = np.zeros((3, nu)) + 0.1
    cpu.funk_loop_dev(2, -1, 0.001, 3, 0.1, 0.01, 0.02, u, v, ri, ra)
    assert np.array_equal(rec.svd_u, u) and np.array_equal(rec.svd_v, v)

    # feature_training_dev: estimator_loop with batch 0 and a max_epochs * dim history, returned
    seen = {}

    def fake_loop(mn, mx, mi, dim, fi, lr, K, u, v, ri, ra, batch, hist, nbu, nbi, verbose=0):
        seen.update(batch=batch, hist_len=hist.shape[0], nbu=nbu, nbi=nbi)
        seen_order["ri"], seen_order["ra"] = ri.copy(), ra.copy()
        cpu.funk_loop_dev(mn, mx, mi, dim, fi, lr, K, u, v, ri, ra, batch, hist)

    monkeypatch.setattr(gd_estimator, "estimator_loop", fake_loop)
    seen_order = {}
    np.random.seed(5)
    rec = make()
    hist = rec.feature_training_dev()
    assert seen == {"batch": 0, "hist_len": 5 * 3, "nbu": nu, "nbi": ni}
    # the reference shuffles (RNG consumed) and then overwrites the arrays with ratings_iterator()
    # order (gradient_descent.py:588-592): the loop trains on the UNSHUFFLED lil -> coo order
    ri0, ra0 = rec.get_ratings()
    assert np.array_equal(seen_order["ri"], ri0) and np.array_equal(seen_order["ra"], ra0)
    np.random.seed(5)
    np.random.shuffle(np.arange(n))
    state_after_one_shuffle = np.random.get_state()[1][:4].copy()
    np.random.seed(5)
    make().feature_training_dev()
    assert np.array_equal(np.random.get_state()[1][:4], state_after_one_shuffle)   # exactly one shuffle drawn
    assert hist.shape == (15,) and (hist.reshape(3, 5)[:, :2] > 0).all()

    # feature_training_bias: bias statistics first, then the learned-bias loop with K -> K_feature, K2 -> K_bias
    def fake_lb(mn, mx, mi, dim, fi, lr, lru, lri, Kf, Kb, mu, u, v, ri, ra, ib, ub, nbu, nbi, verbose=0):
        seen.update(Kf=Kf, Kb=Kb, mu=mu, ib0=ib.copy())
        cpu.funk_learned_bias(mn, mi, dim, fi, lr, lru, lri, Kf, Kb, mu, u, v, ri, ra, ib, ub)

    monkeypatch.setattr(gd_estimator, "estimator_loop_with_learned_bias", fake_lb)
    from mfrec_b200 import _native
    monkeypatch.setattr(_native, "bias_stats", lambda idx_, r_, ni_, nu_, K2=0.01, K3=0.01, ctx=None:
                        cpu.bias_stats(idx_, r_, ni_, nu_, K2, K3))
    np.random.seed(5)
    rec = make()
    rec.feature_training_bias()
    assert seen["Kf"] == rec.K and seen["Kb"] == rec.K2 and abs(seen["mu"] - r.mean()) < 1e-12
    assert seen["ib0"].any() and not np.array_equal(seen["ib0"], rec.items_bias)   # biases were learned in place
