#!/usr/bin/env python
"""CPU simulation of hot-item splitting with the float64 oracle as the kernel:
items with more than `tau` ratings are trained as J = ceil(d / tau) independent copies (each sees the
ratings of the users with hash(user) mod J == j) that are merged after every epoch.  Prints the
train / probe RMSE next to plain sequential training so the merge rule can be judged before it
goes into the packer.  TEST/DESIGN TOOL: imports oracle/, never used by the product.

    python tools/hot_split_sim.py c3p 8400 mean
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mfrec_b200 import synth  # noqa: E402
from oracle import cpu  # noqa: E402

spec = importlib.util.spec_from_file_location("mc", os.path.join(ROOT, "tests", "golden", "make_convergence.py"))
mc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mc)


def main():
    name, tau, rule = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    p = mc.problem(name)
    idx, r, nu, ni, k = p["idx"], p["r"], p["nu"], p["ni"], p["k"]
    deg = np.bincount(idx[:, 1], minlength=ni)
    J = np.maximum(1, -(-deg // tau))
    vbase = np.concatenate([[0], np.cumsum(J)])
    ni_v = int(vbase[-1])
    h = (idx[:, 0].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
    copy = (h % J[idx[:, 1]].astype(np.uint64)).astype(np.int64)
    vidx = idx.copy()
    vidx[:, 1] = (vbase[idx[:, 1]] + copy).astype(np.int32)
    item_of = np.repeat(np.arange(ni), J)
    print("%s: %d items, %d split into %d copies in total (tau %d, hottest %d ratings), rule %s"
          % (name, ni, int((J > 1).sum()), int(J[J > 1].sum()), tau, deg.max(), rule), flush=True)
    u, v = synth.init_factors(nu, ni, k, seed=2)
    ib, ub = np.zeros(ni), np.zeros(nu)
    uv = np.ascontiguousarray(u[:, item_of])
    ibv = ib[item_of].copy()
    hot = np.nonzero(J > 1)[0]
    for e in range(p["epochs"]):
        u_prev, ib_prev = u.copy(), ib.copy()
        rm = cpu.kmf_train("linear", 1, k, p["lr"], mc.K_USERS, mc.K_ITEMS, mc.K_BIAS, uv, v, vidx, r, ibv, ub)
        # merge
        u[:, J == 1] = uv[:, vbase[:-1][J == 1]]
        ib[J == 1] = ibv[vbase[:-1][J == 1]]
        for i in hot:
            cols = slice(vbase[i], vbase[i + 1])
            if rule == "mean":
                u[:, i] = uv[:, cols].mean(axis=1)
                ib[i] = ibv[cols].mean()
            else:   # sum of deltas
                u[:, i] = u_prev[:, i] + (uv[:, cols] - u_prev[:, [i]]).sum(axis=1)
                ib[i] = ib_prev[i] + (ibv[cols] - ib_prev[i]).sum()
        uv = np.ascontiguousarray(u[:, item_of])
        ibv = ib[item_of].copy()
        print("epoch %d running rmse %.6f" % (e + 1, rm[0]), flush=True)
    tr = mc.rmse_numpy("linear", u, v, ib, ub, idx, r)
    pr = mc.rmse_numpy("linear", u, v, ib, ub, p["probe_idx"], p["probe_r"])
    print("split: train %.6f probe %.6f" % (tr, pr), flush=True)


if __name__ == "__main__":
    main()
