"""The reference's import paths resolve to this implementation (no GPU needed to import):
kmf.py:15-18, gradient_descent.py:20-25, mf.py:19-23 of the reference."""
import numpy as np
import pytest


def test_reference_import_lines_work_unchanged():
    from mfrec.lib.datasets import create_bool_sparse_row, create_bool_sparse_col  # noqa: F401
    from mfrec.recommendation.base import BaseRecommender  # noqa: F401
    from mfrec.recommendation.mf import MFRecommender  # noqa: F401
    from mfrec.lib.machinelearning.kmf_train import train_logistic_kernel, train_linear_kernel  # noqa: F401  (kmf.py:18)
    from mfrec.recommendation.metrics import test_predict_rating  # noqa: F401
    from mfrec.lib.gd_estimator import estimator_loop, estimator_loop2, estimator_loop_with_bias, \
        estimator_loop_with_bias_dev, estimator_subloop, predictor_subloop, estimator_loop_without_bias, \
        estimator_loop_with_implicit_feedback, estimator_loop_with_learned_bias  # noqa: F401
    from mfrec.lib.kmf_train import train_linear_kernel as t2
    from mfrec.recommendation.kmf import KMFRecommender
    from mfrec.recommendation.gradient_descent import GDRecommender  # noqa: F401
    from mfrec.recommendation.wrmf import WRMFRecommender  # noqa: F401
    from mfrec.lib.als_implicit import als_wrmf  # noqa: F401
    import mfrec.lib.kmf_train
    import mfrec_b200.lib.kmf_train
    assert mfrec.lib.kmf_train is mfrec_b200.lib.kmf_train       # the same module object: shared state
    assert t2 is train_linear_kernel
    rec = KMFRecommender(5, 7)
    assert rec.nbr_users == 5 and rec.nbr_items == 7


def test_short_arrays_raise_before_the_native_call():
    """every gd_estimator entry point checks what kmf_train checks (no GPU is touched:
    the checks run first)."""
    from mfrec.lib import gd_estimator, kmf_train
    u, v = np.zeros((3, 6)) + 0.1, np.zeros((3, 4)) + 0.1
    idx = np.zeros((5, 2), dtype=np.int32)
    r = np.ones(8)                                           # more ratings than index rows
    with pytest.raises(ValueError):
        gd_estimator.estimator_loop_without_bias(2, 2, 1e-4, 3, 0.1, 0.01, 0.05, u, v, idx, r, 4, 6)
    with pytest.raises(ValueError):
        gd_estimator.estimator_loop(2, 2, 1e-4, 3, 0.1, 0.01, 0.05, u, v, idx, r, 0, np.zeros(6), 4, 3)
    with pytest.raises(ValueError):                          # bias arrays shorter than the factor arrays
        gd_estimator.estimator_loop_with_bias(2, 2, 1e-4, 3, 0.1, 0.01, 0.0, 0.0, 0.05, 3.0, u, v, idx, np.ones(5),
                                              np.zeros(2), np.zeros(4), 4, 6)
    with pytest.raises(ValueError):
        gd_estimator.estimator_loop_with_learned_bias(2, 2, 1e-4, 3, 0.1, 0.01, 0.01, 0.01, 0.05, 0.01, 3.0, u, v, idx,
                                                      np.ones(5), np.zeros(6), np.zeros(3), 4, 6)
    with pytest.raises(ValueError):                          # index array of the wrong width
        kmf_train.train_linear_kernel(1, 3, 0.1, 0.01, 0, 0, 0.1, 0.1, 0.007, 0.0, u, v,
                                      np.zeros((5, 3), dtype=np.int32), np.ones(5), np.zeros(6), np.zeros(4))
