"""ctypes binding of ``mfrec_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

Every function keeps the reference's array conventions: ``u`` = item factors ``[k, ni]``,
``v`` = user factors ``[k, nu]``, both float64 C-contiguous and mutated in place;
``ratings_index`` int32 ``[nnz, 2]`` = (user, item); ``ratings`` float64 ``[nnz]``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "libmfrec_oracle.so")
_lib = None

PREDICTORS = {
    "predict_rating": 0,            # gradient_descent.py:621
    "predict_rating_with_bias": 1,  # gradient_descent.py:637
    "predict_linear": 2,            # kmf.py:88
    "predict_logistic": 3,          # kmf.py:79
    "predict_linear_neg": 4,        # kmf.py:97
    "predict_dot": 5,               # wrmf.py:67
}


def build(force=False):
    src = os.path.join(_DIR, "mfrec_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "_build/libmfrec_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_funk_train.restype = C.c_int64
        _lib.oracle_topn_user.restype = C.c_int
        _lib.oracle_bias_stats.restype = C.c_double
    return _lib


def _p(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def _zeros_like_bias(n, b):
    return np.zeros(n, dtype=np.float64) if b is None else b


def kmf_train(kernel, nbr_epochs, dim, lr, K_users, K_items, K_bias, u, v, ratings_index,
              ratings, items_bias, users_bias, update_users=1, update_items=1):
    """kernel: 'linear' | 'logistic'.  Returns rmse per epoch (float64[nbr_epochs])."""
    kid = {"linear": 0, "logistic": 1}[kernel]
    rm = np.zeros(max(nbr_epochs, 1), dtype=np.float64)
    lib().oracle_kmf_train(
        C.c_int(kid), C.c_int(nbr_epochs), C.c_int(dim), C.c_double(lr), C.c_double(K_users),
        C.c_double(K_items), C.c_double(K_bias), _p(u, np.float64), _p(v, np.float64),
        _p(ratings_index, np.int32), _p(ratings, np.float64), C.c_int64(ratings.shape[0]),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), _p(items_bias, np.float64),
        _p(users_bias, np.float64), C.c_int(update_users), C.c_int(update_items),
        _p(rm, np.float64))
    return rm[:nbr_epochs]


def funk_train(variant, min_epochs, min_improvement, dim, f_init, lr, K, u, v, ratings_index,
               ratings, overall_avg=1.0, items_bias=None, users_bias=None, update_users=1,
               update_items=1):
    """variant: 'without_bias' | 'with_bias' | 'with_bias_dev'.
    Returns (passes, epochs_per_feature, rmse_per_feature)."""
    vid = {"without_bias": 0, "with_bias": 1, "with_bias_dev": 2}[variant]
    ib = _zeros_like_bias(u.shape[1], items_bias)
    ub = _zeros_like_bias(v.shape[1], users_bias)
    fe = np.zeros(dim, dtype=np.int32)
    fr = np.zeros(dim, dtype=np.float64)
    passes = lib().oracle_funk_train(
        C.c_int(vid), C.c_int(min_epochs), C.c_double(min_improvement), C.c_int(dim),
        C.c_double(f_init), C.c_double(lr), C.c_double(K), C.c_double(overall_avg),
        _p(u, np.float64), _p(v, np.float64), _p(ratings_index, np.int32),
        _p(ratings, np.float64), C.c_int64(ratings.shape[0]), C.c_int64(u.shape[1]),
        C.c_int64(v.shape[1]), _p(ib, np.float64), _p(ub, np.float64), C.c_int(update_users),
        C.c_int(update_items), _p(fe, np.int32), _p(fr, np.float64))
    return int(passes), fe, fr


def kmf_epoch_rowmajor_f32(dim, lr, K_users, K_items, K_bias, P, Q, users_bias, items_bias, ratings_index,
                           ratings):
    """One epoch of the linear kernel's arithmetic on row-major float32 factors (measurement aid for
    bench.py's fair-layout CPU figure); returns the sum of squared errors."""
    fn = lib().oracle_kmf_epoch_rowmajor_f32
    fn.restype = C.c_double
    return float(fn(C.c_int(dim), C.c_float(lr), C.c_float(K_users), C.c_float(K_items), C.c_float(K_bias),
                    _p(P, np.float32), _p(Q, np.float32), _p(users_bias, np.float32), _p(items_bias, np.float32),
                    _p(ratings_index, np.int32), _p(ratings, np.float32), C.c_int64(ratings.shape[0])))


def kmf_epoch_rowmajor_f32_mt(dim, lr, K_users, K_items, K_bias, P, Q, users_bias, items_bias, ratings_index,
                              ratings, threads):
    """The same loop on `threads` host threads (user slices; item rows shared without locks): a
    throughput figure for bench.py's fair-layout CPU line, nothing is compared with its result."""
    fn = lib().oracle_kmf_epoch_rowmajor_f32_mt
    fn.restype = C.c_double
    return float(fn(C.c_int(dim), C.c_float(lr), C.c_float(K_users), C.c_float(K_items), C.c_float(K_bias),
                    _p(P, np.float32), _p(Q, np.float32), _p(users_bias, np.float32), _p(items_bias, np.float32),
                    _p(ratings_index, np.int32), _p(ratings, np.float32), C.c_int64(ratings.shape[0]),
                    C.c_int32(P.shape[0]), C.c_int(int(threads))))


def funk_loop_dev(min_epochs, max_epochs, min_improvement, dim, f_init, lr, K, u, v, ratings_index,
                  ratings, batch=0, rmse_hist=None):
    """estimator_loop (max_epochs >= 0, rmse_hist required) / estimator_loop2 (max_epochs < 0).
    Returns (epochs_per_feature, rmse_per_feature)."""
    fe = np.zeros(dim, dtype=np.int32)
    fr = np.zeros(dim, dtype=np.float64)
    lib().oracle_funk_loop_dev(
        C.c_int(min_epochs), C.c_int(max_epochs), C.c_double(min_improvement), C.c_int(dim),
        C.c_double(f_init), C.c_double(lr), C.c_double(K), _p(u, np.float64), _p(v, np.float64),
        _p(ratings_index, np.int32), _p(ratings, np.float64), C.c_int64(ratings.shape[0]),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), C.c_int(batch),
        _p(rmse_hist, np.float64) if rmse_hist is not None else None, _p(fe, np.int32), _p(fr, np.float64))
    return fe, fr


def funk_subloop(f, dim, f_init, lr, K, u, v, ratings_index, ratings, rating_cache):
    fn = lib().oracle_funk_subloop
    fn.restype = C.c_double
    return float(fn(C.c_int(f), C.c_int(dim), C.c_double(f_init), C.c_double(lr), C.c_double(K),
                    _p(u, np.float64), _p(v, np.float64), _p(ratings_index, np.int32),
                    _p(ratings, np.float64), C.c_int64(ratings.shape[0]), C.c_int64(u.shape[1]),
                    C.c_int64(v.shape[1]), _p(rating_cache, np.float64)))


def funk_predictor_subloop(f, dim, f_init, u, v, ratings_index, rating_cache):
    lib().oracle_funk_predictor_subloop(
        C.c_int(f), C.c_int(dim), C.c_double(f_init), _p(u, np.float64), _p(v, np.float64),
        _p(ratings_index, np.int32), C.c_int64(ratings_index.shape[0]), C.c_int64(u.shape[1]),
        C.c_int64(v.shape[1]), _p(rating_cache, np.float64))


def funk_learned_bias(min_epochs, min_improvement, dim, f_init, lr, lr_users, lr_items, K_feature,
                      K_bias, overall_avg, u, v, ratings_index, ratings, items_bias, users_bias):
    fe = np.zeros(dim, dtype=np.int32)
    fr = np.zeros(dim, dtype=np.float64)
    lib().oracle_funk_learned_bias(
        C.c_int(min_epochs), C.c_double(min_improvement), C.c_int(dim), C.c_double(f_init),
        C.c_double(lr), C.c_double(lr_users), C.c_double(lr_items), C.c_double(K_feature),
        C.c_double(K_bias), C.c_double(overall_avg), _p(u, np.float64), _p(v, np.float64),
        _p(ratings_index, np.int32), _p(ratings, np.float64), C.c_int64(ratings.shape[0]),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), _p(items_bias, np.float64),
        _p(users_bias, np.float64), _p(fe, np.int32), _p(fr, np.float64))
    return fe, fr


def _bias_args(u, v, mu, items_bias, users_bias):
    ib = _zeros_like_bias(u.shape[1], items_bias)
    ub = _zeros_like_bias(v.shape[1], users_bias)
    return C.c_double(mu), _p(ib, np.float64), _p(ub, np.float64), ib, ub


def predict_pairs(predictor, u, v, pairs, mu=0.0, items_bias=None, users_bias=None,
                  min_rating=1.0, max_rating=5.0):
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    out = np.zeros(pairs.shape[0], dtype=np.float64)
    cmu, pib, pub, _ib, _ub = _bias_args(u, v, mu, items_bias, users_bias)
    lib().oracle_predict_pairs(
        C.c_int(PREDICTORS[predictor]), _p(u, np.float64), _p(v, np.float64),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), C.c_int(u.shape[0]),
        _p(pairs, np.int32), C.c_int64(pairs.shape[0]), cmu, pib, pub,
        C.c_double(min_rating), C.c_double(max_rating), _p(out, np.float64))
    return out


def rmse_pairs(predictor, u, v, pairs, real, mu=0.0, items_bias=None, users_bias=None,
               min_rating=1.0, max_rating=5.0):
    """Returns (stats[4] = rmse, mae, var_abs, n_valid; errors[n])."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    real = np.ascontiguousarray(real, dtype=np.float64)
    errs = np.zeros(pairs.shape[0], dtype=np.float64)
    out = np.zeros(4, dtype=np.float64)
    cmu, pib, pub, _ib, _ub = _bias_args(u, v, mu, items_bias, users_bias)
    lib().oracle_rmse_pairs(
        C.c_int(PREDICTORS[predictor]), _p(u, np.float64), _p(v, np.float64),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), C.c_int(u.shape[0]),
        _p(pairs, np.int32), _p(real, np.float64), C.c_int64(pairs.shape[0]), cmu, pib, pub,
        C.c_double(min_rating), C.c_double(max_rating), _p(errs, np.float64),
        _p(out, np.float64))
    return out, errs


def topn_user(predictor, u, v, user, n_candidates, rated_items, N, mu=0.0, items_bias=None,
              users_bias=None, min_rating=1.0, max_rating=5.0):
    rated = np.ascontiguousarray(rated_items, dtype=np.int32)
    items = np.zeros(max(N, 1), dtype=np.int32)
    scores = np.zeros(max(N, 1), dtype=np.float64)
    cmu, pib, pub, _ib, _ub = _bias_args(u, v, mu, items_bias, users_bias)
    m = lib().oracle_topn_user(
        C.c_int(PREDICTORS[predictor]), _p(u, np.float64), _p(v, np.float64),
        C.c_int64(u.shape[1]), C.c_int64(v.shape[1]), C.c_int(u.shape[0]), C.c_int(user),
        C.c_int(n_candidates), _p(rated, np.int32), C.c_int64(rated.shape[0]), cmu, pib, pub,
        C.c_double(min_rating), C.c_double(max_rating), C.c_int(N), _p(items, np.int32),
        _p(scores, np.float64))
    return items[:m], scores[:m]


def bias_stats(ratings_index, ratings, ni, nu, K2=0.01, K3=0.01):
    """Returns (mu, items_bias, users_bias) -- base.py:504-508, mf.py:78-121."""
    ib = np.zeros(ni, dtype=np.float64)
    ub = np.zeros(nu, dtype=np.float64)
    mu = lib().oracle_bias_stats(
        _p(np.ascontiguousarray(ratings_index, dtype=np.int32), np.int32),
        _p(np.ascontiguousarray(ratings, dtype=np.float64), np.float64),
        C.c_int64(ratings.shape[0]), C.c_int64(ni), C.c_int64(nu), C.c_double(K2),
        C.c_double(K3), _p(ib, np.float64), _p(ub, np.float64))
    return float(mu), ib, ub


def als_wrmf(nbr_epochs, dim, u, v, users_row, users_col, items_row, items_col, c_pos=1, k=0.015):
    """als_implicit.pyx:208-352; in place on u [dim, ni], v [dim, nu]."""
    ur, uc = np.ascontiguousarray(users_row, np.int32), np.ascontiguousarray(users_col, np.int32)
    ir, ic = np.ascontiguousarray(items_row, np.int32), np.ascontiguousarray(items_col, np.int32)
    lib().oracle_als_wrmf(C.c_int(nbr_epochs), C.c_int(dim), _p(u, np.float64), _p(v, np.float64),
                          _p(ur, np.int32), C.c_int64(ur.shape[0]), _p(uc, np.int32), _p(ir, np.int32),
                          C.c_int64(ir.shape[0]), _p(ic, np.int32), C.c_int64(v.shape[1]),
                          C.c_int64(u.shape[1]), C.c_int(c_pos), C.c_double(k))
