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
        rec.similar_items(1, method='euclidean')
