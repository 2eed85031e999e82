#!/usr/bin/env python
"""End-of-training RMSE of THE REFERENCE'S OWN KERNELS (oracle/_ref = the unmodified
mfrec/lib/kmf_train.pyx) on the named configurations, written to tests/golden/convergence.json.

    python tests/golden/make_convergence.py [c1 c2 c3p ...]      (only where /root/reference exists)

The GPU box has no reference checkout, so the handful of floats are committed; the `-m gpu`
tests (tests/test_convergence_gpu.py) regenerate the SAME seeded ratings with mfrec_b200.synth,
train them through the drop-in `train_linear_kernel` / `train_logistic_kernel` and compare.

Configurations (hyper-parameters lr = 0.005 (C1: 0.01), K_users = K_items = 0.05,
K_bias = 0.007, init N(0, 0.1) seed 2, 90/10 train/probe split, order shuffled once):
  c1   MovieLens-100K shape, k = 20, 30 epochs                       (BASELINE configs[0])
  c2   MovieLens-20M shape, k = 64, 10 epochs                        (BASELINE configs[1])
  c3p  Netflix shape 480k x 17.7k, k = 128, a 10M-rating sample with full-size factor matrices
       (BASELINE.md section 3 allows a prefix: a full 100M epoch is ~5 min on one core), 5 epochs
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mfrec_b200 import synth  # noqa: E402

CONFIGS = {
    # name: (shape key, nnz override, k, epochs, lr)
    "c1": ("ml100k", None, 20, 30, 0.01),
    "c2": ("ml20m", None, 64, 10, 0.005),
    "c3p": ("netflix", 10_000_000, 128, 5, 0.005),
}
K_USERS = K_ITEMS = 0.05
K_BIAS = 0.007
OUT = os.environ.get("MFREC_CONVERGENCE_OUT", os.path.join(HERE, "convergence.json"))


def problem(name):
    """The seeded ratings of a configuration (shared with the GPU test)."""
    shape, nnz_override, k, epochs, lr = CONFIGS[name]
    nu, ni, nnz, _ = synth.SHAPES[shape]
    if nnz_override:
        nnz = nnz_override
    d = synth.make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3, probe_frac=0.1)
    d.update(nu=nu, ni=ni, k=k, epochs=epochs, lr=lr)
    return d


def rmse_numpy(kernel, u, v, ib, ub, idx, r):
    """kmf.py:79-94 predictors + metrics.py:69-73, vectorised in float64."""
    se, step = 0.0, 1 << 21
    for a in range(0, idx.shape[0], step):
        us, it = idx[a:a + step, 0], idx[a:a + step, 1]
        s = np.einsum("kn,kn->n", u[:, it], v[:, us]) + ib[it] + ub[us]
        if kernel == "logistic":
            s = 1.0 + 4.0 / (1.0 + np.exp(-s))
        e = r[a:a + step] - s
        se += float(np.dot(e, e))
    return float(np.sqrt(se / idx.shape[0]))


def main():
    from oracle import ref
    kmf = ref.kmf_train()
    names = sys.argv[1:] or list(CONFIGS)
    res = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            res = json.load(f)
    for name in names:
        p = problem(name)
        for kernel, fn in (("linear", kmf.train_linear_kernel), ("logistic", kmf.train_logistic_kernel)):
            u, v = synth.init_factors(p["nu"], p["ni"], p["k"], seed=2)
            ib, ub = np.zeros(p["ni"]), np.zeros(p["nu"])
            t0 = time.time()
            fn(p["epochs"], p["k"], 0.1, p["lr"], 0.0, 0.0, K_USERS, K_ITEMS, K_BIAS, 0.0, u, v,
               p["idx"], p["r"], ib, ub, 1, 1, 0)
            dt = time.time() - t0
            res["%s_%s" % (name, kernel)] = {
                "train_rmse": rmse_numpy(kernel, u, v, ib, ub, p["idx"], p["r"]),
                "probe_rmse": rmse_numpy(kernel, u, v, ib, ub, p["probe_idx"], p["probe_r"]),
                "nnz_train": int(p["idx"].shape[0]), "nnz_probe": int(p["probe_idx"].shape[0]),
                "epochs": p["epochs"], "k": p["k"], "lr": p["lr"], "K_users": K_USERS, "K_items": K_ITEMS,
                "K_bias": K_BIAS, "reference_seconds": round(dt, 1),
                "made_by": "oracle/_ref kmf_train.%s (unmodified mfrec/lib/kmf_train.pyx)" % fn.__name__}
            print(name, kernel, res["%s_%s" % (name, kernel)], flush=True)
            with open(OUT, "w") as f:
                json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
