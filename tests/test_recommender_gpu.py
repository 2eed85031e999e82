"""The recommender classes end to end on the GPU (reference: kmf.py:197-220,
gradient_descent.py:506-545, metrics.py:51-130, mf.py:144-193), checked against the CPU oracle
replaying the reference's own driver steps with the same numpy RNG state."""
import numpy as np
import pytest

from mfrec_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def sequential_schedule():
    from mfrec_b200.lib._buffers import options
    old = options["schedule"]
    options["schedule"] = "sequential"
    yield
    options["schedule"] = old


def _fill(rec, nu, ni, nnz, seed=0):
    d = synth.make_ratings(nu, ni, nnz, seed=seed, shuffle_seed=None)
    rec.set_ratings(d["idx"], d["r"])
    return d


def test_kmf_train_reference_order_is_bit_exact(sequential_schedule):
    from mfrec_b200.recommendation import KMFRecommender
    from oracle import cpu
    nu, ni, nnz = 80, 50, 900
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 4, 'nbr_features': 8, 'learning_rate': 0.02})
    _fill(rec, nu, ni, nnz)
    np.random.seed(5)
    rec.train(kernel='train_linear_kernel')
    # the reference's driver, step by step (kmf.py:207-220), with the oracle as the kernel
    np.random.seed(5)
    u = np.random.normal(0.0, 0.1, [8, ni])
    v = np.random.normal(0.0, 0.1, [8, nu])
    idx, r = rec.get_ratings(randomize_order=True)
    ib, ub = np.zeros(ni), np.zeros(nu)
    cpu.kmf_train("linear", 4, 8, 0.02, 0.1, 0.1, 0.007, u, v, idx, r, ib, ub)
    assert np.array_equal(rec.svd_u, u) and np.array_equal(rec.svd_v, v)
    assert np.array_equal(rec.items_bias, ib) and np.array_equal(rec.users_bias, ub)
    assert abs(rec.overall_bias - r.mean()) < 1e-12


def test_gd_train_reference_order_is_bit_exact(sequential_schedule):
    from mfrec_b200.recommendation import GDRecommender
    from oracle import cpu
    nu, ni, nnz = 70, 40, 800
    params = {'min_epochs': 3, 'max_epochs': 3, 'nbr_features': 5, 'learning_rate': 0.01}
    for handle_bias in (False, True):
        rec = GDRecommender(nu, ni, params)
        _fill(rec, nu, ni, nnz, seed=2)
        np.random.seed(9)
        rec.train(handle_bias=handle_bias)
        np.random.seed(9)
        u = np.zeros([5, ni]) + 0.1
        v = np.zeros([5, nu]) + 0.1
        idx, r = rec.get_ratings(randomize_order=True)
        if handle_bias:
            mu, ib, ub = cpu.bias_stats(idx, r, ni, nu, 0.01, 0.01)
            cpu.funk_train("with_bias", 3, 0.0001, 5, 0.1, 0.01, 0.05, u, v, idx, r, r.mean(), ib, ub)
            np.testing.assert_allclose(rec.items_bias, ib, rtol=1e-10, atol=1e-12)
        else:
            cpu.funk_train("without_bias", 3, 0.0001, 5, 0.1, 0.01, 0.05, u, v, idx, r)
        if handle_bias:   # biases come from a device reduction: equal to ~1e-12, not bit for bit
            np.testing.assert_allclose(rec.svd_u, u, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(rec.svd_v, v, rtol=1e-9, atol=1e-11)
        else:
            assert np.array_equal(rec.svd_u, u) and np.array_equal(rec.svd_v, v)


def test_metrics_rmse_matches_the_python_loop(capsys):
    from mfrec_b200.recommendation import KMFRecommender, metrics
    nu, ni, nnz = 200, 120, 5000
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 6, 'nbr_features': 16})
    d = _fill(rec, nu, ni, nnz)
    np.random.seed(1)
    rec.train(kernel='train_linear_kernel')
    u_test = np.c_[d["idx"][:700], d["r"][:700]]
    for predictor in ('predict_linear', 'predict_logistic', 'predict'):
        rmse, errors = metrics.test_predict_rating(rec, u_test, nbr_samples=500, predictor=predictor)
        fn = getattr(rec, predictor)
        want = np.array([row[2] - fn(int(row[1]), int(row[0])) for row in u_test[:500]])
        np.testing.assert_allclose(errors, want, rtol=1e-5, atol=1e-5)
        assert abs(rmse - np.sqrt((want ** 2).mean())) <= 1e-5 * rmse
    assert 'Mean root mean square error (RMSE)' in capsys.readouterr().out


def _reference_top(rec, user_index, n_candidates, nbr, predictor):
    """mf.py:156-190 / gradient_descent.py:776-802 restated literally."""
    already = set(rec._rated_items(user_index).tolist()) | {user_index}
    scores = np.zeros(n_candidates)
    for i in range(n_candidates):
        scores[i] = 0.0 if i in already else getattr(rec, predictor)(i, user_index)
    scores[np.isnan(scores)] = 0.0
    nz = scores.nonzero()[0]
    order = sorted(nz, key=lambda i: scores[i], reverse=True)[:nbr]
    return [int(i) for i in order], [scores[i] for i in order]


def test_top_n_entry_points_keep_the_reference_quirks():
    from mfrec_b200.recommendation import GDRecommender, KMFRecommender, metrics
    nu, ni, nnz = 90, 140, 3000
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 5, 'nbr_features': 12})
    d = _fill(rec, nu, ni, nnz)
    np.random.seed(2)
    rec.train()
    rec.neighborhood = 100
    for user in (0, 17, 89):
        items, scores = rec.find_recommended_items(user_index=user, nbr_recommendations=7)
        want_items, want_scores = _reference_top(rec, user, 100, 7, 'predict')
        assert user not in items and max(items) < 100
        np.testing.assert_allclose(scores, want_scores, rtol=1e-5)
        assert items == want_items
    labels, _ = rec.find_recommended_items(user_label='user3', nbr_recommendations=3, output_label=True)
    assert all(lab.startswith('item') for lab in labels)
    p, r, f = metrics.precision_recall(rec, np.c_[d["idx"][:300], d["r"][:300]], nbr_recommendations=5)
    assert 0.0 <= p <= 1.0 and 0.0 <= r <= 1.0

    gd = GDRecommender(nu, ni, {'min_epochs': 2, 'max_epochs': 2, 'nbr_features': 6})
    gd.set_ratings(d["idx"], d["r"])
    np.random.seed(3)
    gd.train()
    items, scores = gd.find_user_top_match(11, nbr_recommendations=9)
    want_items, want_scores = _reference_top(gd, 11, ni, 9, 'predict_rating')
    np.testing.assert_allclose(scores, want_scores, rtol=1e-5)
    assert items == want_items


def test_fold_in_a_new_user():
    from mfrec_b200.recommendation import KMFRecommender
    nu, ni, nnz = 60, 40, 1500
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 8, 'nbr_features': 8})
    _fill(rec, nu, ni, nnz)
    np.random.seed(4)
    rec.train(kernel='train_linear_kernel')
    items_before = rec.svd_u.copy()
    new_id = rec.add_user('alice', np.array([1, 5, 9, 20]), np.array([5.0, 4.0, 1.0, 3.0]))
    assert new_id == nu and rec.svd_v.shape[1] == nu + 1 and rec.users_index['alice'] == nu
    assert np.array_equal(rec.svd_u, items_before)            # items stay frozen (update_items = 0)
    assert np.isfinite(rec.predict_linear(5, new_id))


def test_similar_items_match_the_reference_loops():
    """base.py:1420-1466 / gradient_descent.py:827-875: similarity of one item to every item, sorted,
    the item itself dropped."""
    from mfrec_b200.recommendation import GDRecommender, KMFRecommender
    rng = np.random.default_rng(0)
    nu, ni, k = 30, 90, 12
    rec = KMFRecommender(nu, ni, {'nbr_features': k})
    rec.svd_u = rng.normal(0, 0.3, (k, ni))
    rec.svd_v = rng.normal(0, 0.3, (k, nu))

    def reference(rows, q, n, method):
        sims = []
        for coord in rows:
            a, b = coord, rows[q]
            if method == 'pearson':
                a, b = a - a.mean(), b - b.mean()
            ip = np.inner(a, b)
            sims.append(ip / (np.linalg.norm(a) * np.linalg.norm(b)) if ip != 0 else 0.0)
        order = sorted(range(len(sims)), key=lambda i: sims[i], reverse=True)
        return order[1:n + 1], [sims[i] for i in order[1:n + 1]]

    for method in ('cosine', 'pearson'):
        ids, sims = rec.similar_items(17, nbr_recommendations=6, similarities_output=True, method=method)
        wi, ws = reference(rec.svd_u.T, 17, 6, method)
        assert ids == wi
        np.testing.assert_allclose(sims, ws, rtol=1e-5, atol=1e-6)
    assert rec.similar_items(3, nbr_recommendations=4) == reference(rec.svd_u.T, 3, 4, 'cosine')[0]
    labels = rec.similar_items_by_label('item5', nbr_recommendations=3)
    assert labels == ['item%d' % i for i in reference(rec.svd_u.T, 5, 3, 'cosine')[0]]
    high = rec.similar_items(17, nbr_recommendations='All', similarity_threshold=0.3, similarities_output=True)
    assert all(v > 0.3 for v in high[1]) and 17 not in high[0]
    gd = GDRecommender(nu, ni, {'nbr_features': k})
    gd.svd_u, gd.svd_v = rec.svd_u.copy(), rec.svd_v.copy()
    ids, sims = gd.similar_items(8, nbr_recommendations=5, similarities_output=True)      # Pearson on features 1..k-1
    wi, ws = reference(gd.svd_u[1:k, :].T, 8, 5, 'pearson')
    assert ids == wi
    np.testing.assert_allclose(sims, ws, rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        rec.similar_items(1, method='manhattan')

    # euclidean: the reference sorts DISTANCES in descending order and drops the head of the list
    # (base.py:1434-1466), i.e. it returns the farthest items but one; similar_users removes the
    # user itself explicitly and keeps the head (base.py:1335)
    def reference_euclid(rows, q, n, drop_first):
        d = [float(np.linalg.norm(c - rows[q])) for c in rows]
        order = sorted(range(len(d)), key=lambda i: d[i], reverse=True)
        if not drop_first:
            order = [i for i in order if i != q]
        order = order[1:n + 1] if drop_first else order[:n]
        return order, [d[i] for i in order]

    ids, dist = rec.similar_items(17, nbr_recommendations=5, similarities_output=True, method='euclidean')
    wi, wd = reference_euclid(rec.svd_u.T, 17, 5, True)
    assert ids == wi
    np.testing.assert_allclose(dist, wd, rtol=1e-4)
    ids, dist = rec.similar_users(4, nbr_recommendations=5, similarities_output=True, method='euclidean')
    wi, wd = reference_euclid(rec.svd_v.T, 4, 5, False)
    assert ids == wi
    np.testing.assert_allclose(dist, wd, rtol=1e-4)
    ids, sims = rec.similar_users(9, nbr_recommendations=4, similarities_output=True)     # Pearson by default
    wi, ws = reference(rec.svd_v.T, 9, 4, 'pearson')
    assert ids == wi
    np.testing.assert_allclose(sims, ws, rtol=1e-5, atol=1e-6)
    # many queries in one call == one query at a time; the prepared rows are uploaded once
    bi, bs, bc = rec.similar_items_batch(np.arange(ni), nbr_recommendations=6)
    for q in (0, 17, 89):
        assert [int(i) for i in bi[q, :bc[q]]] == rec.similar_items(q, nbr_recommendations=6)


def test_resident_model_is_reused_until_the_factors_change():
    """predict / RMSE / top-N through the reference's entry points keep the factors in HBM
    (the model is uploaded once, not per call) and notice retraining."""
    from mfrec_b200 import _native
    from mfrec_b200.recommendation import KMFRecommender, metrics
    nu, ni, nnz = 300, 200, 6000
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 3, 'nbr_features': 16})
    d = _fill(rec, nu, ni, nnz)
    np.random.seed(1)
    rec.train(kernel='train_linear_kernel')
    u_test = np.c_[d["idx"][:50], d["r"][:50]]
    m1 = rec._resident_model()
    metrics.test_predict_rating(rec, u_test, nbr_samples=10, predictor='predict_linear')
    rec.find_recommended_items(user_index=3, nbr_recommendations=5)
    assert rec._resident_model() is m1                       # same device copy, nothing re-uploaded
    np.random.seed(2)
    rec.train(kernel='train_linear_kernel')                 # new factors (new arrays)
    m2 = rec._resident_model()
    assert m2 is not m1
    rmse, errors = metrics.test_predict_rating(rec, u_test, nbr_samples=10, predictor='predict_linear')
    want = np.array([row[2] - rec.predict_linear(int(row[1]), int(row[0])) for row in u_test[:10]])
    np.testing.assert_allclose(errors, want, rtol=1e-5, atol=1e-5)
    rec.svd_u[:, 7] += 1.0                                    # an element-wise edit: tell the cache
    rec.invalidate_model()
    assert rec._resident_model() is not m2


def test_precision_recall_is_one_sweep_for_all_users():
    """metrics.precision_recall (metrics.py:85-130) over >= 10k test users: ONE batched top-N call
    (launch count independent of the number of users), same numbers as the per-user loop."""
    from mfrec_b200 import _native
    from mfrec_b200.recommendation import KMFRecommender, metrics
    rng = np.random.default_rng(3)
    nu, ni, k = 12000, 2500, 32
    d = synth.make_ratings(nu, ni, 200000, seed=4, shuffle_seed=None)
    rec = KMFRecommender(nu, ni, {'nbr_features': k})
    rec.set_ratings(d["idx"], d["r"])
    rec.svd_u = rng.normal(0, 0.3, (k, ni))
    rec.svd_v = rng.normal(0, 0.3, (k, nu))
    rec.items_bias, rec.users_bias = rng.normal(0, 0.1, ni), rng.normal(0, 0.1, nu)
    rec.overall_bias = 3.5
    rec.relationship_matrix_csc = rec.relationship_matrix.T.tocsc()
    rec.neighborhood = ni
    held = np.c_[rng.integers(0, nu, 40000), rng.integers(0, ni, 40000), np.ones(40000)]
    n_users = np.unique(held[:, 0]).shape[0]
    assert n_users >= 10000
    ctx = _native.default_context()
    rec._resident_model()
    l0 = ctx.launch_count
    p, r, f = metrics.precision_recall(rec, held, nbr_recommendations=10)
    launches = ctx.launch_count - l0
    assert launches < 200, launches                          # one sweep launch set, not one per user
    # the same numbers from the reference's per-user loop on a subset of the users
    some = held[np.isin(held[:, 0], np.unique(held[:, 0])[:40])]
    items, _s, counts = rec.find_recommended_items_batch(np.unique(some[:, 0]).astype(np.int32), 10)
    for j, user in enumerate(np.unique(some[:, 0]).astype(int)):
        one, _ = rec.find_recommended_items(user_index=user, nbr_recommendations=10)
        assert set(one) == set(int(i) for i in items[j, :counts[j]])   # (fp32 summation order may swap near-ties)
    assert 0.0 <= p <= 1.0 and 0.0 <= r <= 1.0


def test_few_pairs_do_not_upload_the_model():
    """test_predict_rating's default is 10 samples: with Netflix-sized factors the call must stay
    under a millisecond-scale budget (resident model), and the one-shot C entry points gather only
    the rows the pairs name."""
    import time
    from mfrec_b200 import _native
    from mfrec_b200.recommendation import KMFRecommender, metrics
    from oracle import cpu
    rng = np.random.default_rng(5)
    nu, ni, k = synth.SHAPES["netflix"][0], synth.SHAPES["netflix"][1], 128
    rec = KMFRecommender(4, 6, {'nbr_features': k})
    rec.svd_u = rng.normal(0, 0.1, (k, ni))
    rec.svd_v = rng.standard_normal((k, nu)) * 0.1
    rec.items_bias, rec.users_bias = np.zeros(ni), np.zeros(nu)
    rec.overall_bias = 0.0
    pairs = np.c_[rng.integers(0, nu, 10), rng.integers(0, ni, 10)].astype(np.int32)
    real = rng.integers(1, 6, 10).astype(np.float64)
    u_test = np.c_[pairs, real]
    # one-shot C ABI: compact model
    t0 = time.perf_counter()
    stats, errs = _native.rmse_pairs("predict_linear", rec.svd_u, rec.svd_v, pairs, real, 0.0, rec.items_bias, rec.users_bias)
    dt_oneshot = time.perf_counter() - t0
    want, werr = cpu.rmse_pairs("predict_linear", rec.svd_u, rec.svd_v, pairs, real, 0.0, rec.items_bias, rec.users_bias)
    np.testing.assert_allclose(errs, werr, rtol=1e-5, atol=1e-6)
    assert abs(stats[0] - want[0]) <= 1e-5 * want[0]
    assert dt_oneshot < 0.05, dt_oneshot                     # was a 0.5 GB upload (~40 ms of PCIe alone + conversion)
    # recommender entry point: resident model, first call uploads, later calls do not
    metrics.test_predict_rating(rec, u_test, nbr_samples=10, predictor='predict_linear')
    best = 1e9
    for _ in range(20):
        t0 = time.perf_counter()
        _native_model = rec._resident_model()
        pred, _ = _native_model.predict("predict_linear", pairs, None, 0.0, 1.0, 5.0)
        best = min(best, time.perf_counter() - t0)
    np.testing.assert_allclose(real - pred, werr, rtol=1e-5, atol=1e-6)
    print("10 pairs on resident Netflix-shaped factors: %.3f ms per call" % (best * 1e3))
    assert best < 1e-3, best


def test_fold_in_leaves_the_frozen_side_bit_identical():
    """retrain_user (kmf.py:120-131, update_items = 0): svd_u, items... stay exactly as they were,
    and every other user's row too (a stratified fp32 pass rounded them in place)."""
    from mfrec_b200.recommendation import KMFRecommender
    from oracle import cpu
    nu, ni, nnz = 60, 40, 1500
    rec = KMFRecommender(nu, ni, {'nbr_epochs': 8, 'nbr_features': 8})
    d = _fill(rec, nu, ni, nnz)
    np.random.seed(4)
    rec.train(kernel='train_linear_kernel')
    # give the model non-fp32-representable values: any fp32 round trip would change them
    rec.svd_u += 1e-12
    rec.svd_v += 1e-12
    u0, v0, ib0, ub0 = rec.svd_u.copy(), rec.svd_v.copy(), rec.items_bias.copy(), rec.users_bias.copy()
    idx, r = d["idx"], d["r"]
    np.random.seed(6)
    rec.retrain_user(7, idx, r, kernel='train_linear_kernel')
    assert np.array_equal(rec.svd_u, u0)
    others = np.arange(nu) != 7
    assert np.array_equal(rec.svd_v[:, others], v0[:, others])
    assert np.array_equal(rec.users_bias[others], ub0[others])
    # and the retrained row is the reference's own result (sequential order, float64)
    np.random.seed(6)
    v_ref, ib_ref, ub_ref = v0.copy(), ib0.copy(), ub0.copy()
    v_ref[:, 7] = np.random.normal(0.0, 0.1, 8)
    m = idx[:, 0] == 7
    u_ref = u0.copy()
    cpu.kmf_train("linear", 8, 8, rec.learning_rate, rec.K_users, rec.K_items, rec.K_bias, u_ref, v_ref,
                  np.ascontiguousarray(idx[m]), np.ascontiguousarray(r[m]), ib_ref, ub_ref, 1, 0)
    assert np.array_equal(rec.svd_v[:, 7], v_ref[:, 7])
    assert np.array_equal(rec.items_bias, ib_ref)
