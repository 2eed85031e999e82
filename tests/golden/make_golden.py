#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN KERNELS (oracle/_ref, built
unmodified from /root/reference/mfrec/lib/*.pyx by oracle/build_ref.sh) on small seeded inputs.

The reference ships no tests or fixtures; these files are the pinned
known-answer vectors for the C oracle (tests/test_oracle.py) and for the CUDA sequential
schedule (tests/test_golden_gpu.py).  Re-run only where /root/reference exists:

    python tests/golden/make_golden.py

Predictor / RMSE vectors use the reference's one-line numpy formulas verbatim
(gradient_descent.py:629,645-646; kmf.py:83-85,92-94,101-103; metrics.py:69-73), because the
Python-2 classes that hold them cannot be imported under Python 3.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mfrec_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402


def problem(nu=50, ni=40, nnz=600, seed=11):
    d = synth.make_ratings(nu, ni, nnz, seed=seed, shuffle_seed=seed + 1)
    return nu, ni, d["idx"], d["r"]


def main():
    kmf, gd = ref.kmf_train(), ref.gd_estimator()
    nu, ni, idx, r = problem()
    k = 8

    # ---- A1: train_linear_kernel / train_logistic_kernel (kmf_train.pyx:195-277, 103-189)
    for name, fn in (("linear", kmf.train_linear_kernel), ("logistic", kmf.train_logistic_kernel)):
        for gates in ((1, 1), (1, 0), (0, 1)):
            u, v = synth.init_factors(nu, ni, k, seed=5)
            ib, ub = np.zeros(ni), np.zeros(nu)
            u0, v0 = u.copy(), v.copy()
            fn(4, k, 0.1, 0.02, 0.5, 0.5, 0.05, 0.07, 0.007, 3.3, u, v, idx, r, ib, ub, gates[0], gates[1], 0)
            np.savez_compressed(
                os.path.join(HERE, "kmf_%s_u%d_i%d.npz" % (name, gates[0], gates[1])),
                idx=idx, r=r, u0=u0, v0=v0, u=u, v=v, ib=ib, ub=ub, nbr_epochs=4, k=k, lr=0.02,
                K_users=0.05, K_items=0.07, K_bias=0.007, update_users=gates[0], update_items=gates[1])

    # ---- A2: Funk-SVD loops (gd_estimator.pyx:691-779, 489-582, 588-685)
    order = np.lexsort((idx[:, 1], idx[:, 0]))
    mu = float(r.mean())
    K2 = K3 = 0.01
    bi = np.zeros(ni)
    bu = np.zeros(nu)
    for i in range(ni):   # mf.py:78-97 compute_items_bias_bk
        m = idx[order, 1] == i
        if m.any():
            bi[i] = (r[order][m] - mu).sum() / (K3 + m.sum())
    for j in range(nu):   # mf.py:100-121 compute_users_bias_bk
        m = idx[order, 0] == j
        if m.any():
            bu[j] = (r[order][m] - mu - bi[idx[order, 1][m]]).sum() / (K2 + m.sum())
    f_init, lr, K = 0.1, 0.002, 0.05
    kf = 5
    u = np.zeros((kf, ni)) + f_init
    v = np.zeros((kf, nu)) + f_init
    gd.estimator_loop_without_bias(6, 99, 0.0005, kf, f_init, lr, K, u, v, idx, r, nu, ni, 0)
    np.savez_compressed(os.path.join(HERE, "funk_without_bias.npz"), idx=idx, r=r, u=u, v=v, k=kf,
                        f_init=f_init, lr=lr, K=K, min_epochs=6, min_improvement=0.0005)
    u = np.zeros((kf, ni)) + f_init
    v = np.zeros((kf, nu)) + f_init
    gd.estimator_loop_with_bias(6, 99, 0.0005, kf, f_init, lr, 0.3, 0.3, K, mu, u, v, idx, r, bi, bu, nu, ni, 0)
    np.savez_compressed(os.path.join(HERE, "funk_with_bias.npz"), idx=idx, r=r, u=u, v=v, k=kf,
                        f_init=f_init, lr=lr, K=K, min_epochs=6, min_improvement=0.0005, mu=mu, bi=bi, bu=bu)
    for gates in ((1, 0), (0, 1)):
        u, v = synth.init_factors(nu, ni, kf, seed=6)
        u, v = np.abs(u) + 0.05, np.abs(v) + 0.05
        u0, v0 = u.copy(), v.copy()
        gd.estimator_loop_with_bias_dev(4, 99, 0.0005, kf, f_init, lr, 0.3, 0.3, K, mu, u, v, idx, r, bi, bu,
                                        nu, ni, gates[0], gates[1], 0)
        np.savez_compressed(os.path.join(HERE, "funk_with_bias_dev_u%d_i%d.npz" % gates), idx=idx, r=r,
                            u0=u0, v0=v0, u=u, v=v, k=kf, f_init=f_init, lr=lr, K=K, min_epochs=4,
                            min_improvement=0.0005, mu=mu, bi=bi, bu=bu, update_users=gates[0],
                            update_items=gates[1])

    # ---- A3: development variants (gd_estimator.pyx:210-303, 308-395, 903-995, 401-483)
    kd = 3
    hist = np.zeros(2 * 7 * kd)
    u = np.zeros((kd, ni)) + f_init
    v = np.zeros((kd, nu)) + f_init
    gd.estimator_loop(3, 7, 0.0005, kd, f_init, lr, K, u, v, idx, r, 1, hist, nu, ni, 0)
    out = dict(idx=idx, r=r, k=kd, f_init=f_init, lr=lr, K=K, loop_u=u, loop_v=v, loop_hist=hist,
               loop_min_epochs=3, loop_max_epochs=7, loop_min_improvement=0.0005, loop_batch=1)
    u = np.zeros((kd, ni)) + f_init
    v = np.zeros((kd, nu)) + f_init
    gd.estimator_loop2(4, 99, 0.0005, kd, f_init, lr, K, u, v, idx, r, np.zeros(ni), nu, ni, 0)
    out.update(loop2_u=u, loop2_v=v, loop2_min_epochs=4, loop2_min_improvement=0.0005)
    u = np.zeros((kd, ni)) + f_init
    v = np.zeros((kd, nu)) + f_init
    cache = np.zeros(nu * ni)
    rmses = []
    for f in range(2):
        for _ in range(3):
            rmses.append(gd.estimator_subloop(f, 1, 0.0, kd, f_init, lr, K, u, v, idx, r, cache, nu, ni, 0))
        gd.predictor_subloop(f, 1, kd, f_init, u, v, idx, r, cache, nu, ni)
    cells = idx[:, 0].astype(np.int64) + idx[:, 1].astype(np.int64) * nu
    out.update(sub_u=u, sub_v=v, sub_rmse=np.array(rmses), sub_cache_cells=cells, sub_cache_values=cache[cells],
               sub_cache_sum=cache.sum())
    u = np.zeros((kd, ni)) + f_init
    v = np.zeros((kd, nu)) + f_init
    lb, lub = bi.copy(), bu.copy()
    gd.estimator_loop_with_learned_bias(3, 99, 0.0005, kd, f_init, lr, 0.004, 0.003, K, 0.01, mu, u, v, idx, r,
                                        lb, lub, nu, ni, 0)
    out.update(lb_u=u, lb_v=v, lb_ib0=bi, lb_ub0=bu, lb_ib=lb, lb_ub=lub, lb_mu=mu, lb_min_epochs=3,
               lb_min_improvement=0.0005, lb_lr_users=0.004, lb_lr_items=0.003, lb_K_bias=0.01)
    np.savez_compressed(os.path.join(HERE, "funk_dev.npz"), **out)

    # ---- predictors + RMSE (the reference's numpy one-liners)
    u, v = synth.init_factors(nu, ni, k, seed=7)
    rng = np.random.Generator(np.random.PCG64(8))
    ib, ub = rng.normal(0, 0.2, ni), rng.normal(0, 0.2, nu)
    pairs = idx[:200]
    real = r[:200].copy()
    real[17] = np.nan        # a NaN error must be dropped (metrics.py:70-71)
    mn, mx = 1.0, 5.0
    preds = {}
    for name in ("predict_rating", "predict_rating_with_bias", "predict_linear", "predict_logistic",
                 "predict_linear_neg", "predict_dot"):
        out = np.zeros(pairs.shape[0])
        for j, (user, item) in enumerate(pairs):
            s = np.dot(u[:, item], v[:, user])
            if name == "predict_rating":
                out[j] = s + 1.0
            elif name == "predict_rating_with_bias":
                s += mu + (ib[item] + ub[user])
                out[j] = s
            elif name == "predict_linear":
                s += (ib[item] + ub[user])
                out[j] = s
            elif name == "predict_logistic":
                s += (ib[item] + ub[user])
                out[j] = mn + (1.0 / (1.0 + np.exp(-s))) * (mx - mn)
            elif name == "predict_linear_neg":
                s += (ib[item] + ub[user])
                out[j] = mn + s * (mx - mn)
            else:
                out[j] = s
        preds[name] = out
    stats = {}
    for name, p in preds.items():
        all_errors = real - p
        errors = all_errors[np.where(np.isnan(all_errors) == False)[0]]  # noqa: E712
        abs_errors = abs(errors)
        stats[name] = np.array([np.sqrt(pow(abs(errors), 2).mean()), abs_errors.mean(), abs_errors.var(),
                                float(len(abs_errors))])
    np.savez_compressed(os.path.join(HERE, "predictors.npz"), u=u, v=v, ib=ib, ub=ub, mu=mu, pairs=pairs,
                        real=real, **{"pred_" + n: p for n, p in preds.items()},
                        **{"stats_" + n: s for n, s in stats.items()})

    # ---- top-N (gradient_descent.py:769-802 over all items; mf.py:144-193 over the first M ids)
    rated = {}
    for (user, item) in idx:
        rated.setdefault(int(user), []).append(int(item))
    users = np.array([0, 3, 7, 39, 41], dtype=np.int32)
    N = 6
    out = {}
    for tag, predictor, ncand in (("gd_all", "predict_rating", ni), ("mf_first25", "predict_logistic", 25)):
        items_out = np.full((len(users), N), -1, dtype=np.int32)
        scores_out = np.zeros((len(users), N))
        for row, user in enumerate(users):
            user_ratings = np.zeros(ncand)
            already = np.r_[np.array(sorted(rated.get(int(user), [])), dtype=np.int64), user]
            for i in range(ncand):
                if i not in already:
                    user_ratings[i] = preds_fn(predictor, u, v, ib, ub, mu, i, int(user))
            user_ratings[np.where(np.isnan(user_ratings))] = 0.0
            top = {int(i): user_ratings[i] for i in user_ratings.nonzero()[0]}
            srt = sorted(top.items(), key=lambda kv: kv[1], reverse=True)[:N]
            for c, (i, s) in enumerate(srt):
                items_out[row, c], scores_out[row, c] = i, s
        out["items_" + tag], out["scores_" + tag] = items_out, scores_out
    indptr = np.zeros(len(users) + 1, dtype=np.int64)
    flat = []
    for row, user in enumerate(users):
        flat += sorted(rated.get(int(user), []))
        indptr[row + 1] = len(flat)
    np.savez_compressed(os.path.join(HERE, "topn.npz"), u=u, v=v, ib=ib, ub=ub, mu=mu, users=users,
                        rated_indptr=indptr, rated_items=np.array(flat, dtype=np.int32), N=N, **out)
    print("golden vectors written to", HERE)


def preds_fn(name, u, v, ib, ub, mu, item, user):
    s = np.dot(u[:, item], v[:, user])
    if name == "predict_rating":
        return s + 1.0
    s += (ib[item] + ub[user])
    return 1.0 + (1.0 / (1.0 + np.exp(-s))) * 4.0


if __name__ == "__main__":
    main()
