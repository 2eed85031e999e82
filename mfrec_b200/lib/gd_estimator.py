"""``mfrec.lib.gd_estimator`` on a B200: the Funk-SVD per-feature SGD loops.

Signatures follow mfrec/lib/gd_estimator.pyx: ``estimator_loop_without_bias`` (:691-704),
``estimator_loop_with_bias`` (:489-507), ``estimator_loop_with_bias_dev`` (:588-608).  ``u`` and
``v`` are trained in place, biases are read-only, ``max_epochs`` / ``learning_rate_users`` /
``learning_rate_items`` / ``nbr_users`` / ``nbr_items`` are accepted and unused exactly like the
reference; returns ``None``.  The development-only variants that allocate a dense
users x items cache (``estimator_loop``, ``estimator_loop2``, ``estimator_subloop``,
``predictor_subloop``), the learned-bias hybrid and the (buggy) implicit-feedback loop are out
of scope (priority A3) and raise NotImplementedError.
"""
import numpy as np

from mfrec_b200 import _native
from mfrec_b200.lib._buffers import buffer_arg, native_opts, options

last_feature_epochs = None
last_feature_rmse = None


def _train(variant, min_epochs, max_epochs, min_improvement, dim, f_init, learning_rate, K,
           overall_avg, u, v, ratings_index, ratings, items_bias, users_bias, update_users,
           update_items, verbose):
    global last_feature_epochs, last_feature_rmse
    dim = int(dim)
    buffer_arg(u, "u", np.float64, 2)
    buffer_arg(v, "v", np.float64, 2)
    buffer_arg(ratings_index, "ratings_index", np.int32, 2, writable=False)
    buffer_arg(ratings, "ratings", np.float64, 1, writable=False)
    if variant != _native.FUNK_WITHOUT_BIAS:
        buffer_arg(items_bias, "items_bias", np.float64, 1, writable=False)
        buffer_arg(users_bias, "users_bias", np.float64, 1, writable=False)
    if dim > u.shape[0] or dim > v.shape[0] or dim < 0:
        raise ValueError("dim=%d exceeds the factor arrays (%d, %d features)" % (dim, u.shape[0], v.shape[0]))
    if dim == 0:
        return None
    ctx = _native.default_context(options["device"])
    fe, fr = _native.train_funk(
        variant, int(min_epochs), int(max_epochs), float(min_improvement), dim, float(f_init),
        float(learning_rate), float(K), float(overall_avg), u[:dim], v[:dim], ratings_index,
        ratings, items_bias, users_bias, 1 if update_users else 0, 1 if update_items else 0,
        ctx=ctx, **native_opts())
    last_feature_epochs, last_feature_rmse = fe, fr
    if verbose:
        for f in range(dim):
            print("Training the feature " + str(f))
            print("RMSE: " + str(fr[f]) + "\n")
    return None


def estimator_loop_without_bias(min_epochs, max_epochs, min_improvement, dim, f_init,
                                learning_rate, K, u, v, ratings_index, ratings, nbr_users,
                                nbr_items, verbose=0):
    return _train(_native.FUNK_WITHOUT_BIAS, min_epochs, max_epochs, min_improvement, dim, f_init,
                  learning_rate, K, 1.0, u, v, ratings_index, ratings, None, None, 1, 1, verbose)


def estimator_loop_with_bias(min_epochs, max_epochs, min_improvement, dim, f_init, learning_rate,
                             learning_rate_users, learning_rate_items, K, overall_avg, u, v,
                             ratings_index, ratings, items_bias, users_bias, nbr_users, nbr_items,
                             verbose=0):
    return _train(_native.FUNK_WITH_BIAS, min_epochs, max_epochs, min_improvement, dim, f_init,
                  learning_rate, K, overall_avg, u, v, ratings_index, ratings, items_bias,
                  users_bias, 1, 1, verbose)


def estimator_loop_with_bias_dev(min_epochs, max_epochs, min_improvement, dim, f_init,
                                 learning_rate, learning_rate_users, learning_rate_items, K,
                                 overall_avg, u, v, ratings_index, ratings, items_bias, users_bias,
                                 nbr_users, nbr_items, update_users=1, update_items=1, verbose=0):
    return _train(_native.FUNK_WITH_BIAS_DEV, min_epochs, max_epochs, min_improvement, dim, f_init,
                  learning_rate, K, overall_avg, u, v, ratings_index, ratings, items_bias,
                  users_bias, update_users, update_items, verbose)


def _out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(
            "%s is a development-only variant of the reference (dense users x items cache or "
            "known-buggy loop) and is not part of the B200 hot path; see DESIGN.md" % name)
    fn.__name__ = name
    return fn


estimator_loop = _out_of_scope("estimator_loop")
estimator_loop2 = _out_of_scope("estimator_loop2")
estimator_loop_with_learned_bias = _out_of_scope("estimator_loop_with_learned_bias")
estimator_loop_with_implicit_feedback = _out_of_scope("estimator_loop_with_implicit_feedback")
estimator_subloop = _out_of_scope("estimator_subloop")
predictor_subloop = _out_of_scope("predictor_subloop")
