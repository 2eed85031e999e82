import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mfrec_b200 import synth
from mfrec_b200.lib import kmf_train
dev = torch.device('cuda', 0)
nu, ni, nnz, k = synth.SHAPES['netflix']
idx_d, r_d = bench.gpu_synth(torch, dev, nu, ni, nnz, seed=0)
pinned = len(sys.argv) > 1 and sys.argv[1] == 'pinned'
idx_h = torch.empty((nnz, 2), dtype=torch.int32, pin_memory=pinned); idx_h.copy_(idx_d)
r_h = torch.empty(nnz, dtype=torch.float64, pin_memory=pinned); r_h.copy_(r_d.double())
del idx_d, r_d
u0, v0 = synth.init_factors(nu, ni, k, seed=2)
u_h = torch.from_numpy(u0); v_h = torch.from_numpy(v0)
if pinned: u_h, v_h = u_h.pin_memory(), v_h.pin_memory()
ib = np.zeros(ni); ub = np.zeros(nu)
for it in range(int(os.environ.get("CALLS", "3"))):
    t0 = time.perf_counter()
    kmf_train.train_linear_kernel(1, k, 0.1, 0.005, 0.0, 0.0, 0.05, 0.05, 0.007, 0.0, u_h.numpy(), v_h.numpy(), idx_h.numpy(), r_h.numpy(), ib, ub)
    print("call %d: %.1f ms rmse %.5f" % (it, (time.perf_counter() - t0) * 1e3, kmf_train.last_rmse[-1]), flush=True)
