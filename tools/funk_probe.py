"""Throughput of the Funk-SVD per-feature path (estimator_loop_without_bias, gd_estimator.pyx:691-779)
on synthetic ratings: feature-updates/s through the drop-in call (host arrays in, host arrays out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mfrec_b200 import synth
from mfrec_b200.lib import gd_estimator

wl = os.environ.get("WORKLOAD", "ml20m")
nu, ni, nnz, _k = synth.SHAPES[wl]
k = int(os.environ.get("K", "4"))
min_epochs = int(os.environ.get("EPOCHS", "3"))
dev = torch.device("cuda", 0)
idx_d, r_d = bench.gpu_synth(torch, dev, nu, ni, nnz, seed=0)
idx = idx_d.cpu().numpy(); r = r_d.double().cpu().numpy()
del idx_d, r_d
for it in range(2):
    u = np.zeros((k, ni)) + 0.1; v = np.zeros((k, nu)) + 0.1
    t0 = time.perf_counter()
    gd_estimator.estimator_loop_without_bias(min_epochs, min_epochs, 1e9, k, 0.1, 0.001, 0.05, u, v, idx, r, nu, ni, 0)
    dt = time.perf_counter() - t0
    passes = int(gd_estimator.last_feature_epochs.sum())
    print("%s: k=%d, %d training passes over %d ratings in %.3f s -> %.2f G feature-updates/s (call incl. transfers), rmse %.4f"
          % (wl, k, passes, nnz, dt, passes * nnz / dt / 1e9, gd_estimator.last_feature_rmse[-1]), flush=True)
