#!/usr/bin/env python
"""Benchmark of the SGD hot path: rating updates/s on synthetic Netflix-shaped ratings.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU kernels

A *step* is one epoch of ``train_linear_kernel`` semantics (kmf_train.pyx:195-277) over the
whole rating set.  ``value`` times K epochs with everything resident in HBM (CUDA events on the
library's stream); ``e2e`` times K calls of the drop-in ``train_linear_kernel(nbr_epochs=1)``
with host (pinned) numpy buffers, i.e. H2D of ratings + factors, layout, one epoch, D2H of the
factors inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "rating_updates_per_s"
UNIT = "updates/s"
HP = dict(lr=0.005, K_users=0.05, K_items=0.05, K_bias=0.007)


def algorithmic_bytes_per_update(k, elem_bytes=4):
    """read+write P_u and Q_i (4*k*s) + rating triple (12) + both biases r/w (16)."""
    return 4 * k * elem_bytes + 12 + 16


def captured_traffic(kernel, updates_per_launch):
    """DRAM bytes per launch of the dominant kernel from the latest committed `ncu --set full`
    capture (profiles/traffic.json), rescaled to this run's updates per launch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)[kernel]
        return t["dram_bytes_per_launch"] * updates_per_launch / t["updates_per_launch"], t["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region.

    nvidia-smi needs a few hundred ms to produce its first row, so the sampler is started BEFORE the
    warm-up steps (`with ClockSampler(i) as c:` around warm-up + timed region) and the timed region
    is marked with c.begin() / c.end(); summary() uses the rows that fall inside the marks.  When
    the timed region is shorter than the sampling period (multi-GPU runs: a few epochs of ~50 ms)
    it falls back to the rows of the second before c.end() -- warm-up steps of the same kernel on
    the same data, i.e. the same load -- and says so in "window"."""

    def __init__(self, device_index):
        self.device_index = device_index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def __exit__(self, *a):
        if self.t1 is None:
            self.t1 = time.monotonic()
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        t0 = self.t0 if self.t0 is not None else -1.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [r for (t, r) in self.rows if t0 <= t <= t1]
        window = "timed region"
        if not inside:
            inside = [r for (t, r) in self.rows if t1 - 1.0 <= t <= t1 + 0.05]
            window = "last second before the end of the timed region (warm-up steps of the same load; the timed region is shorter than the sampling period)"
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# --------------------------------------------------------------------------------------------
# synthetic data on the GPU (torch is plumbing here: RNG, sort/unique, host pinning)
# --------------------------------------------------------------------------------------------
def gpu_synth(torch, dev, nu, ni, nnz, seed, user_offset=0, item_tiles=1, item_seed=None):
    """Netflix-shaped unique (user, item, rating) triples in shuffled order (see
    mfrec_b200/synth.py for the model; this is the same recipe with torch's generator).
    item_tiles > 1: the ni items are `item_tiles` copies of one item-popularity profile (the
    multi-GPU weak-scaling workload); item_seed fixes that profile across ranks."""
    from mfrec_b200 import synth
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ni_tile = ni // item_tiles
    wu, _ = synth.marginals(nu, ni_tile, seed)
    _, wi = synth.marginals(nu, ni_tile, seed if item_seed is None else item_seed)
    cu = torch.from_numpy(np.cumsum(wu)).to(dev)
    ci = torch.from_numpy(np.cumsum(wi)).to(dev)
    keys = None
    while keys is None or keys.numel() < nnz:
        need = nnz - (0 if keys is None else keys.numel())
        m = int(need * 1.12) + 4096
        us = torch.searchsorted(cu, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=nu - 1)
        it = torch.searchsorted(ci, torch.rand(m, device=dev, dtype=torch.float64, generator=g)).clamp_(max=ni_tile - 1)
        if item_tiles > 1:
            it += torch.randint(0, item_tiles, (m,), device=dev, generator=g) * ni_tile
        new = us * ni + it
        del us, it
        keys = torch.unique(new if keys is None else torch.cat([keys, new]))
        del new
    perm = torch.randperm(keys.numel(), device=dev, generator=g)[:nnz]
    keys = keys[perm]
    del perm
    users = (keys // ni).to(torch.int32)
    items = (keys % ni).to(torch.int32)
    del keys
    # planted rank-16 model
    rank = 16
    bu = torch.randn(nu, device=dev, generator=g) * 0.3
    bi = torch.randn(ni, device=dev, generator=g) * 0.3
    p = torch.randn(nu, rank, device=dev, generator=g) * 0.35
    q = torch.randn(ni, rank, device=dev, generator=g) * 0.35
    r = torch.empty(nnz, device=dev, dtype=torch.float32)
    step = 1 << 24
    for a in range(0, nnz, step):
        ul, il = users[a:a + step].long(), items[a:a + step].long()
        val = 3.6 + bu[ul] + bi[il] + (p[ul] * q[il]).sum(1)
        val += torch.randn(val.shape[0], device=dev, generator=g) * 0.5
        r[a:a + step] = val.round_().clamp_(1.0, 5.0)
    idx = torch.stack([users + user_offset, items], dim=1).contiguous()
    return idx, r


def run_native(args):
    import torch
    import torch.distributed as dist
    from mfrec_b200 import _native, synth
    from mfrec_b200.lib import kmf_train

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nu, ni, nnz, k = synth.SHAPES[args.workload]
    if args.nnz:
        nnz = args.nnz
    if world > 1:
        from mfrec_b200 import dsgd
        return dsgd.bench_multi_gpu(args, rank, world, local, nu, ni, nnz, k, HP, gpu_synth,
                                    ClockSampler, algorithmic_bytes_per_update, measured_peaks)

    t_setup = time.time()
    G = max(1, args.emulate_slabs)
    if G > 1:
        # debug: the per-rank work of a G-GPU weak-scaling ring on ONE GPU (G item slabs processed
        # back to back, no exchange) -- not a bench line
        ni = ni * G
    idx_d, r_d = gpu_synth(torch, dev, nu, ni, nnz, seed=0, item_tiles=G)
    torch.cuda.synchronize()
    ctx = _native.default_context(local)   # the context the drop-in modules use as well: one stream, one pool
    u0, v0 = synth.init_factors(nu, ni, k, seed=2)

    # ---------------- device-resident arm: K epochs, CUDA events on the library stream ------
    R = _native.Ratings(None, None, ni, nu, ctx=ctx, device_ptrs=(idx_d.data_ptr(), r_d.data_ptr()),
                        nnz=nnz, ratings_are_f32=True, k_hint=k, row_blocks=args.row_blocks,
                        workers=args.workers, n_slabs=(G if G > 1 else 0))
    M = _native.Model(k, ni, nu, u0, v0, None, None, layout=R, ctx=ctx)
    layout_desc = ("stratified B=%d W=%d sub-epochs/epoch=%d max_bucket=%d widest_column_block=%d items"
                   % (R.B, R.W, R.launches_per_epoch, R.max_bucket, R.max_cb_items))
    # schedule balance: the serial chain of one epoch = sum over launches of the slowest CTA,
    # a CTA = sum over phases of its fullest bucket; ideal = nnz / (B * W)
    _, cnt = R.offsets()
    c4 = cnt.reshape(R.G, R.B, R.B, R.W, R.W).astype(np.int64)   # [g, rb, cb, w, phase]
    cta = c4.max(axis=3).sum(axis=3)                              # [g, rb, cb]
    rbs = np.arange(R.B)
    crit = sum(int(cta[g, rbs, (rbs + s) % R.B].max()) for g in range(R.G) for s in range(R.B))
    quad_types = R.quad_types()
    balance = {"critical_path_ratings": crit, "ideal": nnz / float(R.B * R.W),
               "efficiency": nnz / float(R.B * R.W) / max(crit, 1)}
    del cnt, c4, cta
    se = torch.zeros(args.steps + args.warmup, device=dev, dtype=torch.float64)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for e in range(args.warmup):
            M.sgd_epoch(R, _native.KERNEL_LINEAR, HP["lr"], HP["K_users"], HP["K_items"], HP["K_bias"],
                        sq_err_ptr=se.data_ptr() + 8 * e)
        ctx.sync()
        launches_warm = ctx.launch_count - launches0
        torch.cuda.synchronize()
        launches1 = ctx.launch_count
        clocks.begin()
        ev0.record(stream)
        for e in range(args.steps):
            M.sgd_epoch(R, _native.KERNEL_LINEAR, HP["lr"], HP["K_users"], HP["K_items"], HP["K_bias"],
                        sq_err_ptr=se.data_ptr() + 8 * (args.warmup + e))
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        clocks.end()
        launches_timed = ctx.launch_count - launches1
    ms = ev0.elapsed_time(ev1)
    ms_per_step = ms / args.steps
    value = nnz * args.steps / (ms * 1e-3)
    rmse_curve = torch.sqrt(se / nnz).cpu().numpy().tolist()

    peak, peak_src = measured_peaks()
    bpu = algorithmic_bytes_per_update(k)
    # dominant kernel = sgd_block_kernel (launches_timed - steps reduce launches); per-launch
    # figures: algorithmic bytes of one launch / its average duration
    n_sgd = launches_timed - args.steps
    achieved = value * bpu / 1e9
    upl = nnz * args.steps / max(n_sgd, 1)
    traffic, traffic_src = captured_traffic("sgd_block_kernel", upl)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "kernel": "sgd_block_kernel", "algorithmic_bytes_per_update": bpu,
                "algorithmic_bytes_per_launch": bpu * upl,
                "updates_per_launch": upl,
                "avg_launch_us": ms * 1e3 / max(n_sgd, 1),
                "note": "algorithmic bytes count the Q_i read+write of every update, but a column block's Q rows "
                        "stay in shared memory for a whole sub-epoch and most P-row sectors hit in the 126 MB L2 "
                        "(see traffic: measured DRAM bytes per launch), so frac can exceed 1; the kernel is "
                        "instruction-issue bound, not DRAM bound (profiles/README.md)"}

    # ---------------- end-to-end arm: the public drop-in call with host buffers -----------------
    e2e = None
    if not args.no_e2e and G == 1:
        del M, R
        idx_h = torch.empty((nnz, 2), dtype=torch.int32, pin_memory=True)
        r_h = torch.empty(nnz, dtype=torch.float64, pin_memory=True)
        idx_h.copy_(idx_d)
        r_h.copy_(r_d.double())
        u_h = torch.from_numpy(u0).pin_memory()
        v_h = torch.from_numpy(v0).pin_memory()
        ib_h = torch.zeros(ni, dtype=torch.float64).pin_memory()
        ub_h = torch.zeros(nu, dtype=torch.float64).pin_memory()
        un, vn, ibn, ubn = u_h.numpy(), v_h.numpy(), ib_h.numpy(), ub_h.numpy()
        idxn, rn = idx_h.numpy(), r_h.numpy()
        kmf_train.options["device"] = local

        def one_call():
            kmf_train.train_linear_kernel(1, k, 0.1, HP["lr"], 0.0, 0.0, HP["K_users"], HP["K_items"],
                                          HP["K_bias"], 0.0, un, vn, idxn, rn, ibn, ubn)
            return kmf_train.last_rmse[-1]

        # warm-up calls: the device's stream-ordered memory pool reaches its steady state after two
        # calls (445 / 76 / 70 / 70 ... ms in a fresh process, tools/e2e_probe.py).  The drop-in
        # uses the same library context as the device-resident arm above (two contexts = two
        # streams sharing one pool made single calls take 0.1 - 2.4 s at random).  ms_per_call
        # lists every timed call.
        e2e_warm = max(args.warmup, 3)
        for _ in range(e2e_warm):
            one_call()
        torch.cuda.synchronize()
        per_call = []
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            tc = time.perf_counter()
            last = one_call()
            per_call.append((time.perf_counter() - tc) * 1e3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = idxn.nbytes + rn.nbytes + un.nbytes + vn.nbytes + ibn.nbytes + ubn.nbytes
        d2h = un.nbytes + vn.nbytes + ibn.nbytes + ubn.nbytes + 8
        e2e = {"value": nnz * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3 / args.e2e_steps,
               "steps": args.e2e_steps, "warmup_calls": e2e_warm, "ms_per_call": per_call,
               "epochs_per_call": 1, "last_rmse": float(last)}

    cpu_baseline = None
    if not args.no_cpu:
        cpu_baseline = time_reference(args.workload, nnz_sample=args.cpu_sample, epochs=1)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s-shaped %dx%d nnz=%d k=%d (BASELINE configs[2])" % (args.workload, nu, ni, nnz, k),
                      "kernel": "train_linear_kernel", "schedule": layout_desc, "balance": balance,
                      "quad_types": quad_types,
                      "l2": "inputs (1.2 GB ratings + 255 MB factors) exceed the 126 MB L2",
                      "hyper": HP},
           "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
           "gpu_launches": int(launches_timed), "clocks": clocks.summary(),
           "rmse_per_epoch": rmse_curve, "setup_s": time.time() - t_setup}
    return out


# --------------------------------------------------------------------------------------------
# secondary bench: BASELINE configs[4], top-N scoring of U.V^T on the tensor cores
#   python bench.py --workload topn [--steps K --warmup W]
# --------------------------------------------------------------------------------------------
def run_topn(args):
    import torch
    from mfrec_b200 import _native, synth
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    nu, ni, _nnz, k = synth.SHAPES["netflix"]
    N = 100
    ctx = _native.default_context(local)   # the context the drop-in modules use as well: one stream, one pool
    u0, v0 = synth.init_factors(nu, ni, k, seed=2)
    u_h, v_h = torch.from_numpy(u0).pin_memory(), torch.from_numpy(v0).pin_memory()
    items = torch.empty((nu, N), dtype=torch.int32).pin_memory()
    scores = torch.empty((nu, N), dtype=torch.float64).pin_memory()
    counts = torch.empty(nu, dtype=torch.int32).pin_memory()
    out = (items.numpy(), scores.numpy(), counts.numpy())

    def call():
        return _native.topn_sweep("predict_rating", u_h.numpy(), v_h.numpy(), None, ni, None, None, N,
                                  ctx=ctx, out=out)[3]

    sweep_ms, finish_ms = [], []
    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            call()
        torch.cuda.synchronize()
        clocks.begin()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = call()
            sweep_ms.append(st[2])
            finish_ms.append(st[7])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks.end()
    flops = 2.0 * nu * ni * k
    sm = float(np.mean(sweep_ms))
    peaks = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks = json.load(f)
    peak = float(peaks.get("bf16_tflops", 1590.0))   # the sweep kernel is timed alone: burst figure
    # spot check against the exact (CUDA-core) path
    users = np.random.default_rng(0).permutation(nu)[:64].astype(np.int32)
    wi, ws, wc = _native.topn("predict_rating", u0, v0, users, ni, None, None, N, ctx=ctx)
    ok = bool(all(np.allclose(out[1][x][:wc[j]], ws[j][:wc[j]], rtol=1e-5) for j, x in enumerate(users)))
    return {"metric": "topn_user_item_scores_per_s", "value": nu * float(ni) / ((sm + float(np.mean(finish_ms))) * 1e-3),
            "unit": "scores/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sm + float(np.mean(finish_ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 filter + f32 exact re-score", "data": "synthetic",
            "config": {"workload": "top-%d of U.V^T, netflix-shaped factors %d users x %d items, k=%d (BASELINE configs[4])" % (N, nu, ni, k),
                       "what": "value = device time of sweep (tcgen05) + finish (exact re-score, rank, certify) kernels"},
            "roofline": {"bound": "tensor", "achieved": flops / (sm * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (sm * 1e-3) / 1e12 / peak, "traffic": None,
                         "kernel": "topn_sweep_kernel", "flops_per_launch": flops / 7.0,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback",
                         "sweep_ms": sm, "finish_ms": float(np.mean(finish_ms))},
            "cpu_baseline": None,
            "e2e": {"value": nu * float(ni) * args.steps / dt, "unit": "scores/s",
                    "h2d_bytes_per_step": u0.nbytes + v0.nbytes, "d2h_bytes_per_step": sum(a.nbytes for a in out),
                    "ms_per_step": dt * 1e3 / args.steps},
            "gpu_launches": None, "users_redone_exactly": float(st[0]), "candidates_per_user": float(st[1]),
            "matches_exact_path_on_64_users": ok, "clocks": clocks.summary()}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU kernel (oracle/_ref, else our C port) on host cores
# --------------------------------------------------------------------------------------------
def time_reference(workload, nnz_sample, epochs=1):
    """Times kmf_train.train_linear_kernel (the reference's own Cython build when available) on a
    bounded sample: the first `nnz_sample` ratings of the workload with FULL-SIZE factor matrices
    (so the strided access pattern is the real one).  Single thread: the reference holds the GIL
    and has no threading."""
    from mfrec_b200 import synth
    from oracle import cpu, ref
    nu, ni, nnz, k = synth.SHAPES[workload]
    n = int(min(nnz_sample, nnz))
    rng = np.random.Generator(np.random.PCG64(0))
    wu, wi = synth.marginals(nu, ni, 0)
    cu, ci = np.cumsum(wu), np.cumsum(wi)
    idx = np.empty((n, 2), dtype=np.int32)
    idx[:, 0] = np.minimum(np.searchsorted(cu, rng.random(n)), nu - 1)
    idx[:, 1] = np.minimum(np.searchsorted(ci, rng.random(n)), ni - 1)
    r = rng.integers(1, 6, n).astype(np.float64)
    u, v = synth.init_factors(nu, ni, k, seed=2)
    ib, ub = np.zeros(ni), np.zeros(nu)
    if ref.available():
        kind = "reference"
        fn = ref.kmf_train().train_linear_kernel
        t0 = time.perf_counter()
        fn(epochs, k, 0.1, HP["lr"], 0.0, 0.0, HP["K_users"], HP["K_items"], HP["K_bias"], 0.0,
           u, v, idx, r, ib, ub, 1, 1, 0)
        dt = time.perf_counter() - t0
    else:
        kind = "port"
        t0 = time.perf_counter()
        cpu.kmf_train("linear", epochs, k, HP["lr"], HP["K_users"], HP["K_items"], HP["K_bias"], u, v, idx, r, ib, ub)
        dt = time.perf_counter() - t0
    # the same arithmetic on ROW-major float32 factors (the GPU path's layout), one core: how much of
    # the gap is the reference's feature-major float64 layout (SURVEY 8(d) "fair-layout CPU figure")
    P = np.ascontiguousarray(v.T, dtype=np.float32)
    Q = np.ascontiguousarray(u.T, dtype=np.float32)
    t0 = time.perf_counter()
    cpu.kmf_epoch_rowmajor_f32(k, HP["lr"], HP["K_users"], HP["K_items"], HP["K_bias"], P, Q,
                               np.zeros(nu, dtype=np.float32), np.zeros(ni, dtype=np.float32), idx,
                               r.astype(np.float32))
    dt_fair = time.perf_counter() - t0
    return {"value": n * epochs / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d-rating sample of the %s workload, full-size factor matrices (%dx%d, k=%d), %d epoch, %.1f s"
                      % (n, workload, nu, ni, k, epochs, dt),
            "host_cores_available": os.cpu_count(),
            "fair_layout": {"value": n / dt_fair, "unit": UNIT, "cores": 1,
                            "what": "same loop in C on row-major float32 factors (oracle/mfrec_oracle.c), same sample"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from mfrec_b200 import synth
    nu, ni, nnz, k = synth.SHAPES[args.workload]
    vals = []
    for _ in range(args.warmup):
        time_reference(args.workload, max(args.cpu_sample // 10, 1000))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(time_reference(args.workload, args.cpu_sample))
    dt = time.perf_counter() - t0
    value = float(np.mean([v["value"] for v in vals]))
    base = dict(vals[-1])
    base["value"] = value
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s-shaped %dx%d k=%d, %d-rating sample per step" % (args.workload, nu, ni, k, args.cpu_sample),
                       "kernel": "train_linear_kernel", "hyper": HP},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="netflix", choices=["ml100k", "ml20m", "netflix", "yahoo", "topn"])
    ap.add_argument("--nnz", type=int, default=0, help="override the number of ratings (debug)")
    ap.add_argument("--row-blocks", type=int, default=0)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=5_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--emulate-slabs", type=int, default=1,
                    help="debug: run one rank's share of a G-GPU ring on one GPU (no exchange)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        print("bench.py: warmup < 3 breaks the timing rules; using 3", file=sys.stderr)
        args.warmup = 3
    if args.workload == "topn":
        out = run_topn(args)
    else:
        out = run_reference(args) if args.impl == "reference" else run_native(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
