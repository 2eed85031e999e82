"""One-off parity run at BASELINE configs[1] scale (synthetic MovieLens-20M shape, k = 64):
the drop-in train_linear_kernel on the GPU (stratified schedule, fp32) against the REFERENCE's own
Cython kernel (oracle/_ref, float64, sequential order) from identical seeds and initial factors.
north_star: end-of-training RMSE within 0.5 % relative.  Too slow for the test-suite (the reference
runs ~0.6 M updates/s); the output is committed under profiles/.

    python tools/parity_at_scale.py [--workload ml20m] [--epochs 3] [--nnz N]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from mfrec_b200 import _native, synth  # noqa: E402
from mfrec_b200.lib import kmf_train  # noqa: E402
from oracle import cpu, ref  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ml20m")
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--nnz", type=int, default=0)
args = ap.parse_args()
nu, ni, nnz, k = synth.SHAPES[args.workload]
if args.nnz:
    nnz = args.nnz
hp = dict(lr=0.005, K_users=0.05, K_items=0.05, K_bias=0.007)
t0 = time.time()
d = synth.make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3, probe_frac=0.1)
idx, r = d["idx"], d["r"]
print("data: %d train / %d probe ratings in %.0f s" % (len(r), len(d["probe_r"]), time.time() - t0), flush=True)
u0, v0 = synth.init_factors(nu, ni, k, seed=2)
ug, vg, ibg, ubg = u0.copy(), v0.copy(), np.zeros(ni), np.zeros(nu)
t0 = time.time()
kmf_train.train_linear_kernel(args.epochs, k, 0.1, hp["lr"], 0.0, 0.0, hp["K_users"], hp["K_items"], hp["K_bias"],
                              0.0, ug, vg, idx, r, ibg, ubg)
t_gpu = time.time() - t0
rm_g = kmf_train.last_rmse
ur, vr, ibr, ubr = u0.copy(), v0.copy(), np.zeros(ni), np.zeros(nu)
t0 = time.time()
if ref.available():
    ref.kmf_train().train_linear_kernel(args.epochs, k, 0.1, hp["lr"], 0.0, 0.0, hp["K_users"], hp["K_items"],
                                        hp["K_bias"], 0.0, ur, vr, idx, r, ibr, ubr, 1, 1, 0)
    kind = "reference (oracle/_ref)"
else:
    cpu.kmf_train("linear", args.epochs, k, hp["lr"], hp["K_users"], hp["K_items"], hp["K_bias"], ur, vr, idx, r, ibr, ubr)
    kind = "port (oracle/mfrec_oracle.c)"
t_cpu = time.time() - t0
out = {"workload": "%s-shaped %dx%d nnz=%d k=%d" % (args.workload, nu, ni, len(r), k), "epochs": args.epochs,
       "cpu_kind": kind, "gpu_seconds_incl_transfers": t_gpu, "cpu_seconds": t_cpu, "gpu_train_rmse_per_epoch": [float(x) for x in rm_g]}
for name, (pi, pr) in (("train_sample", (idx[:2000000], r[:2000000])), ("probe", (d["probe_idx"], d["probe_r"]))):
    sg, _ = _native.rmse_pairs("predict_linear", ug, vg, pi, pr, 0.0, ibg, ubg)
    sr, _ = _native.rmse_pairs("predict_linear", ur, vr, pi, pr, 0.0, ibr, ubr)
    out["rmse_" + name] = {"gpu_factors": float(sg[0]), "reference_factors": float(sr[0]),
                           "relative_difference": float(abs(sg[0] - sr[0]) / sr[0])}
out["within_0.5_percent"] = bool(all(out["rmse_" + n]["relative_difference"] < 5e-3 for n in ("train_sample", "probe")))
print(json.dumps(out, indent=1))
