"""predict_kernel + fused RMSE sums on the resident Netflix-shaped model (bench.py secondary_predict) as a
stand-alone probe: pairs/s over 50 M device-resident training pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
import bench
from mfrec_b200 import _native, synth
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
nu, ni, nnz, k = synth.SHAPES["netflix"]
idx_d, r_d = bench.gpu_synth(torch, dev, nu, ni, nnz, seed=0)
ctx = _native.default_context(0)
u0, v0 = synth.init_factors(nu, ni, k, seed=2)
R = _native.Ratings(None, None, ni, nu, ctx=ctx, device_ptrs=(idx_d.data_ptr(), r_d.data_ptr()), nnz=nnz, ratings_are_f32=True, k_hint=k)
M = _native.Model(k, ni, nu, u0, v0, None, None, layout=R, ctx=ctx)
peak, _ = bench.measured_peaks()
out = bench.secondary_predict(torch, dev, _native, ctx, M, idx_d, r_d, k, peak)
print(json.dumps({a: out[a] for a in ("value", "ms_per_call", "rmse_of_the_pairs")}), flush=True)
os._exit(0)
