"""Evaluation entry points (reference: mfrec/recommendation/metrics.py:51-130).

``test_predict_rating`` keeps its signature and return value ``(rmse, errors)``; the per-pair Python
loop over ``recommender.<predictor>(item, user)`` becomes one batched device call
(``mfrec_rmse_pairs``) whenever the predictor is one of the factor-model predictors."""
import numpy as np

from mfrec_b200 import _native


def test_predict_rating(recommender, u_test, nbr_samples=10, verbose=False, predictor='predict_rating'):
    """RMSE / MAE of ``predictor`` on the first ``nbr_samples`` rows ``(user, item, rating)`` of
    ``u_test``; NaN errors are dropped (metrics.py:69-73)."""
    sample = np.asarray(u_test[0:nbr_samples])
    native = getattr(recommender, 'NATIVE_PREDICTORS', {})
    name = predictor
    if name not in native and hasattr(recommender, '_predict_alias'):
        try:
            name = recommender._predict_alias(predictor)
        except (KeyError, AttributeError):
            name = None
    if name in native and sample.shape[0]:
        pairs = np.ascontiguousarray(sample[:, 0:2], dtype=np.int32)
        real = np.ascontiguousarray(sample[:, 2], dtype=np.float64)
        if hasattr(recommender, '_resident_model'):
            # the recommender keeps its factors in HBM: only the pairs and the predictions move
            pred, _ = recommender._resident_model().predict(
                native[name], pairs, None, recommender.overall_bias or 0.0, recommender.min_rating,
                recommender.max_rating)
            all_errors = real - pred
        else:
            _, all_errors = _native.rmse_pairs(native[name], recommender.svd_u, recommender.svd_v, pairs,
                                               real, recommender.overall_bias or 0.0,
                                               recommender.items_bias, recommender.users_bias,
                                               recommender.min_rating, recommender.max_rating)
        if verbose:
            for i, (e, r) in enumerate(zip(all_errors, real)):
                print('Prediction %d: Predicted = %s, Real = %s' % (i, r - e, r))
    else:   # a predictor this library has no kernel for: the reference's own loop
        fn = getattr(recommender, predictor)
        pred = np.array([fn(int(row[1]), int(row[0])) for row in sample], dtype=np.float64)
        all_errors = (sample[:, 2] if sample.shape[0] else np.zeros(0)) - pred
    errors = all_errors[~np.isnan(all_errors)]
    abs_errors = np.abs(errors)
    rmse = np.sqrt((abs_errors ** 2).mean()) if errors.shape[0] else np.nan
    print('\nNumber of succesful rating: ' + str(len(abs_errors)) + '/' + str(nbr_samples))
    if errors.shape[0]:
        print('Mean abs. error: ' + str(abs_errors.mean()))
        print('Variance of the error: ' + str(abs_errors.var()))
        print('Mean root mean square error (RMSE): ' + str(rmse))
        print('NMAE: ' + str(abs_errors.mean() / 1.6))
        print('MAE: ' + str(abs_errors.mean()) + '\n\n')
    return rmse, errors


test_predict_rating.__test__ = False   # an evaluation metric, not a pytest test


def precision_recall(recommender, u_test, nbr_recommendations=5, predictor='predict', verbose=False):
    """Precision / recall / F of the top-N lists against the held-out ratings (metrics.py:85-130)."""
    test_sample_dict = {}
    for rating in u_test:
        test_sample_dict.setdefault(int(rating[0]), []).append(int(rating[1]))
    precision = recall = 0.0
    users_count = 0
    users = list(test_sample_dict.keys())
    if hasattr(recommender, 'find_recommended_items_batch') and len(users) > 1:
        # one device call for all test users instead of one find_recommended_items call each
        items, _scores, counts = recommender.find_recommended_items_batch(users, nbr_recommendations, predictor)
        recommended = [set(int(i) for i in items[j, :counts[j]]) for j in range(len(users))]
    else:
        recommended = []
        for user_index in users:
            try:
                recommended.append(set(recommender.find_recommended_items(
                    user_index=user_index, nbr_recommendations=nbr_recommendations, output_label=False,
                    predictor=predictor)[0]))
            except KeyError:
                recommended.append(None)
    for user_index, recommended_set in zip(users, recommended):
        held_out = test_sample_dict[user_index]
        if recommended_set is None:
            recommended_set = set()
        else:
            users_count += 1
        intersection = float(len(recommended_set.intersection(held_out)))
        precision += intersection / nbr_recommendations
        recall += intersection / len(held_out)
    precision = precision / users_count
    recall = recall / users_count
    f_measure = 2 * (precision * recall) / (precision + recall) if precision + recall else 0.0
    if verbose:
        print('Precision @ ' + str(nbr_recommendations) + ' : ' + str(precision))
        print('Recall @ ' + str(nbr_recommendations) + ' : ' + str(recall))
        print('F-Measure : ' + str(f_measure))
    return precision, recall, f_measure
