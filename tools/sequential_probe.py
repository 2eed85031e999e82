"""Time of the reference-order (MFREC_SCHED_SEQUENTIAL, float64, bit-exact) schedules on ML-100K-shaped
ratings: kmf_sequential_kernel and funk_sequential_kernel run a window of 32 ratings by dependency level
on one warp (results identical to one thread walking the stream)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mfrec_b200 import _native, synth

nu, ni, nnz, k = synth.SHAPES["ml100k"]
d = synth.make_ratings(nu, ni, nnz, seed=1)
idx, r = d["idx"], d["r"]
epochs = int(os.environ.get("EPOCHS", "10"))
for it in range(2):
    u, v = synth.init_factors(nu, ni, k, seed=2)
    ib, ub = np.zeros(ni), np.zeros(nu)
    t0 = time.perf_counter()
    rm = _native.train_kmf(_native.KERNEL_LINEAR, epochs, k, 0.01, 0.05, 0.05, 0.007, u, v, idx, r, ib, ub,
                           schedule=_native.SCHED_SEQUENTIAL)
    dt = time.perf_counter() - t0
    print("kmf sequential: %d x %d ratings, k=%d: %.3f s -> %.2f M updates/s (%.0f ns per rating), rmse %.4f"
          % (epochs, nnz, k, dt, epochs * nnz / dt / 1e6, dt / (epochs * nnz) * 1e9, rm[-1]), flush=True)
