import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mfrec_b200 import _native, synth
nu, ni, nnz, k = synth.SHAPES['netflix']
nu = int(os.environ.get("NU", nu))
N = 100
u, v = synth.init_factors(nu, ni, k, seed=2)
if os.environ.get('PINNED'):
    import torch
    u = torch.from_numpy(u).pin_memory().numpy(); v = torch.from_numpy(v).pin_memory().numpy()
for it in range(2):
    t0 = time.perf_counter()
    items, scores, counts, stats = _native.topn_sweep("predict_rating", u, v, None, ni, None, None, N)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms total; sweep %.1f ms = %.1f TFLOP/s (useful), fallback users %d, cand/user %.1f, overflow %d, z %.3f"
          % (it, dt * 1e3, stats[2], stats[3] / stats[2] / 1e9, stats[0], stats[1], stats[4], stats[6]), flush=True)
# spot-check 50 users against the exact path
users = np.random.default_rng(0).permutation(nu)[:50].astype(np.int32)
wi, ws, wc = _native.topn("predict_rating", u, v, users, ni, None, None, N)
ok = all(np.allclose(scores[x][:wc[j]], ws[j][:wc[j]], rtol=1e-5) for j, x in enumerate(users))
print("spot check vs exact path:", ok, "min count", counts.min())
