"""Multi-GPU stratified SGD (DSGD) for the KMF kernels: one process per GPU, item-factor slabs
rotating around a ring.

The rating matrix is cut into G user slices (one per rank, resident for the whole run) and G item
slabs.  One epoch = G steps; in step t rank r updates block (users of r) x (slab (r + t) mod G)
with the single-GPU stratified kernel (``mfrec_sgd_epoch(..., slab=c)``), then hands the slab's
item factors and item biases to rank r - 1 and receives the next slab from rank r + 1
(``torch.distributed`` P2P = ncclSend / ncclRecv over NVLink).  No two ranks ever hold the same
slab or the same users, so the update is conflict-free across GPUs exactly as it is across CTAs
and warps inside one GPU.  After G steps every rank holds its own slab again.

The ring logic is backend-neutral (``Ring``): the GPU backend drives libmfrec_b200, the tests
drive the same class over gloo with the CPU oracle as the per-block update.
"""


def slab_at(rank, step, world):
    """Slab a rank works on in a given step of an epoch."""
    return (rank + step) % world


def ring_peers(rank, world):
    """(destination of the slab I just finished, source of the slab I need next)."""
    return (rank - 1) % world, (rank + 1) % world


class Ring(object):
    """Epoch driver.  ``backend`` provides:

    process_slab(c) -> None          update my users x slab c (asynchronous is fine)
    slab_tensors(c) -> [tensor, ..]  the buffers that make up slab c (item factors, item biases);
                                     sent after step t, received for slab (c + 1)
    sq_err() -> float                this rank's sum of squared errors of the finished epoch
    """

    def __init__(self, backend, rank, world, dist=None):
        self.backend, self.rank, self.world, self.dist = backend, rank, world, dist

    def exchange(self, done_slab, next_slab):
        if self.world == 1:
            return
        dist = self.dist
        dst, src = ring_peers(self.rank, self.world)
        stage = getattr(self.backend, "stage", None)
        send = self.backend.slab_tensors(done_slab)
        recv = self.backend.slab_tensors(next_slab)
        if stage:   # communicate through buffers the collective library can map directly
            send, recv, finish = stage(send, recv)
        ops = []
        for t in send:
            ops.append(dist.P2POp(dist.isend, t, dst))
        for t in recv:
            ops.append(dist.P2POp(dist.irecv, t, src))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if stage:
            finish()

    def epoch(self, trace=None):
        """trace: optional callable(label) invoked at the step boundaries (CUDA event timing)."""
        for step in range(self.world):
            c = slab_at(self.rank, step, self.world)
            if trace:
                trace("k%d" % step)
            self.backend.process_slab(c)
            if trace:
                trace("x%d" % step)
            self.exchange(c, slab_at(self.rank, step + 1, self.world))
        if trace:
            trace("end")
        return self.backend.sq_err()

    def gather_items(self):
        """After any whole number of epochs rank r holds slab r: broadcast every slab from its
        owner so all ranks end with the complete item side."""
        if self.world == 1:
            return
        for c in range(self.world):
            for t in self.backend.slab_tensors(c):
                self.dist.broadcast(t, src=c)


# ---------------------------------------------------------------------------------------------
# GPU backend
# ---------------------------------------------------------------------------------------------
class _DevArray(object):
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape),
                                         "typestr": typestr, "version": 2, "strides": None}


class GpuBackend(object):
    def __init__(self, torch, native, ctx, ratings, model, kernel, hp, update_users=1, update_items=1):
        self.torch, self.native, self.ctx = torch, native, ctx
        self.R, self.M, self.kernel, self.hp = ratings, model, kernel, hp
        self.uu, self.ui = update_users, update_items
        (q_ptr, ib_ptr, _p, _ub), (ni, _nu, kpad) = model.device_ptrs()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.Q = torch.as_tensor(_DevArray(q_ptr, (ni, kpad), "<f4"), device=dev)
        self.ib = torch.as_tensor(_DevArray(ib_ptr, (ni,), "<f4"), device=dev)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.se = torch.zeros(ratings.G, device=dev, dtype=torch.float64)
        self.bounds = [ratings.slab_items(c) for c in range(ratings.G)]

    def stage(self, send, recv):
        """The model lives in the library's stream-ordered pool (cudaMallocAsync), which NCCL cannot
        map for direct peer copies; a 9 MB send through such memory took ~2 ms instead of ~50 us.
        Stage both directions through torch-allocated buffers (two device copies of a few MB)."""
        torch = self.torch
        key = tuple(t.shape for t in send) + tuple(t.shape for t in recv)
        if getattr(self, "_stage_key", None) != key:
            self._stage_send = [torch.empty_like(t) for t in send]
            self._stage_recv = [torch.empty_like(t) for t in recv]
            self._stage_key = key
        for s_, t in zip(self._stage_send, send):
            s_.copy_(t)
        bufs_recv = self._stage_recv

        def finish():
            for t, r_ in zip(recv, bufs_recv):
                t.copy_(r_)
        return self._stage_send, bufs_recv, finish

    def process_slab(self, c):
        self.M.sgd_epoch(self.R, self.kernel, self.hp["lr"], self.hp["K_users"], self.hp["K_items"],
                         self.hp["K_bias"], self.uu, self.ui, slab=c,
                         sq_err_ptr=self.se.data_ptr() + 8 * c)

    def slab_tensors(self, c):
        a, b = self.bounds[c]
        return [self.Q[a:b], self.ib[a:b]]

    def sq_err(self):
        with self.torch.cuda.stream(self.stream):
            return self.se.sum()


def bench_multi_gpu(args, rank, world, local, nu, ni, nnz, k, hp, gpu_synth, ClockSampler,
                    bytes_per_update, measured_peaks):
    """bench.py's N > 1 arm: weak scaling, one Netflix-shaped tile per GPU.

    weak  : nu * N users, ni * N items, nnz * N ratings (each rank: nu users, nnz ratings over all
            N * ni items) -- per-GPU work fixed, the regime DSGD is built for
    strong: the single-GPU problem cut in N user slices
    """
    import time

    import torch
    import torch.distributed as dist
    from mfrec_b200 import _native, synth

    dev = torch.device("cuda", local)
    if args.scaling == "weak":
        nu_r, nnz_r, ni_tot = nu, nnz, ni * world
    else:
        nu_r, nnz_r, ni_tot = nu // world, nnz // world, ni
    t_setup = time.time()
    idx_d, r_d = gpu_synth(torch, dev, nu_r, ni_tot, nnz_r, seed=1000 + rank,
                           item_tiles=(world if args.scaling == "weak" else 1), item_seed=0)
    deg = torch.bincount(idx_d[:, 1].long(), minlength=ni_tot)
    dist.all_reduce(deg)
    ctx = _native.Context(local)
    R = _native.Ratings(None, None, ni_tot, nu_r, ctx=ctx,
                        device_ptrs=(idx_d.data_ptr(), r_d.data_ptr()), nnz=nnz_r,
                        ratings_are_f32=True, k_hint=k, n_slabs=world, row_blocks=args.row_blocks,
                        workers=args.workers, item_degree=deg.cpu().numpy())
    u0, v0 = synth.init_factors(nu_r, ni_tot, k, seed=2)      # same item init on every rank
    _, v0 = synth.init_factors(nu_r, 1, k, seed=100 + rank)
    M = _native.Model(k, ni_tot, nu_r, u0, v0, None, None, layout=R, ctx=ctx)
    be = GpuBackend(torch, _native, ctx, R, M, _native.KERNEL_LINEAR, hp)
    ring = Ring(be, rank, world, dist)
    nnz_total = nnz_r * world

    import os
    tracing = bool(os.environ.get("MFREC_DSGD_TRACE"))
    marks = []

    def mark(label):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(be.stream)
        marks.append((label, ev))

    def one_epoch():
        with torch.cuda.stream(be.stream):
            se = ring.epoch(mark if tracing else None)
            dist.all_reduce(se)
        return se

    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ses = []
    with ClockSampler(local) as clocks:   # started before the warm-up: see its docstring
        for _ in range(args.warmup):
            one_epoch()
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        del marks[:]
        launches1 = ctx.launch_count
        clocks.begin()
        ev0.record(be.stream)
        for _ in range(args.steps):
            ses.append(one_epoch())
        ev1.record(be.stream)
        ctx.sync()
        torch.cuda.synchronize()
        dist.barrier()
        clocks.end()
    if tracing:
        import sys
        kern = exch = 0.0
        per_step = {}
        for (la, ea), (lb, eb) in zip(marks[:-1], marks[1:]):
            if la == "end":
                continue
            dt_ = ea.elapsed_time(eb)
            per_step[la] = per_step.get(la, 0.0) + dt_ / args.steps
            if la.startswith("k"):
                kern += dt_
            else:
                exch += dt_
        print("[dsgd trace] rank %d per step: %s" % (rank, " ".join("%s=%.2f" % kv for kv in sorted(per_step.items()))),
              file=sys.stderr, flush=True)
        n_ep = args.steps
        print("[dsgd trace] rank %d: per epoch: slab kernels %.2f ms, exchange (incl. waiting for the neighbour) %.2f ms"
              % (rank, kern / n_ep, exch / n_ep), file=sys.stderr, flush=True)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # device time, max over ranks
    ms = float(ms.item())
    launches_timed = ctx.launch_count - launches1
    value = nnz_total * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    bpu = bytes_per_update(k)
    achieved = value / world * bpu / 1e9
    (a0, b0) = R.slab_items(0)
    out = {"metric": "rating_updates_per_s", "value": value, "unit": "updates/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "%s-shaped tile per GPU: %d users x %d items, nnz=%d, k=%d; total %d x %d, nnz=%d"
                                  % (args.workload, nu_r, ni_tot // (world if args.scaling == "weak" else 1),
                                     nnz_r, k, nu_r * world, ni_tot, nnz_total),
                      "parallelism": "dsgd ring of %d, item slab = %d rows x %d B per hop" % (world, b0 - a0, 4 * (R.B and M.device_ptrs()[1][2]) + 4),
                      "kernel": "train_linear_kernel",
                      "schedule": "stratified B=%d W=%d slabs=%d sub-epochs/epoch=%d" % (R.B, R.W, R.G, R.launches_per_epoch),
                      "l2": "inputs exceed the 126 MB L2", "hyper": hp},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "kernel": "sgd_block_kernel", "per": "GPU", "algorithmic_bytes_per_update": bpu},
           "cpu_baseline": None, "e2e": None,
           "gpu_launches": int(launches_timed), "launches_warmup": int(launches1 - launches0),
           "clocks": clocks.summary(),
           "rmse_per_epoch": [float(torch.sqrt(s / nnz_total).item()) for s in ses],
           "setup_s": time.time() - t_setup}
    # end-to-end: the same epochs including the H2D of this rank's ratings and the layout pass
    if not args.no_e2e:
        idx_h = torch.empty((nnz_r, 2), dtype=torch.int32, pin_memory=True)
        r_h = torch.empty(nnz_r, dtype=torch.float32, pin_memory=True)
        idx_h.copy_(idx_d)
        r_h.copy_(r_d)
        deg_np = deg.cpu().numpy()
        del be, ring, M, R
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        R2 = _native.Ratings(idx_h.numpy(), r_h.numpy(), ni_tot, nu_r, ctx=ctx, k_hint=k, n_slabs=world,
                             row_blocks=args.row_blocks, workers=args.workers, item_degree=deg_np)
        M2 = _native.Model(k, ni_tot, nu_r, u0, v0, None, None, layout=R2, ctx=ctx)
        be2 = GpuBackend(torch, _native, ctx, R2, M2, _native.KERNEL_LINEAR, hp)
        ring2 = Ring(be2, rank, world, dist)
        with torch.cuda.stream(be2.stream):
            for _ in range(args.e2e_steps):
                se = ring2.epoch()
                dist.all_reduce(se)
            ring2.gather_items()
        ctx.sync()
        u1, v1, ib1, ub1 = M2.read()
        last = float(torch.sqrt(se / nnz_total).item())
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": nnz_total * args.e2e_steps / dt, "unit": "updates/s",
                      "h2d_bytes_per_step": (idx_h.numel() * 4 + r_h.numel() * 4 + u0.nbytes + v0.nbytes) // args.e2e_steps,
                      "d2h_bytes_per_step": (u1.nbytes + v1.nbytes + ib1.nbytes + ub1.nbytes + 8) // args.e2e_steps,
                      "ms_per_step": dt * 1e3 / args.e2e_steps, "steps": args.e2e_steps,
                      "what": "pack + %d epochs + gather + read-back in one timed region, per step" % args.e2e_steps,
                      "last_rmse": last}
    dist.barrier()
    dist.destroy_process_group()
    return out if rank == 0 else None
