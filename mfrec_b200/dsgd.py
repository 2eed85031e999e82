"""Multi-GPU stratified SGD (DSGD) for the KMF kernels: one process per GPU, item-factor slabs
rotating around a ring.

Two transports for the same schedule:

* ``PeerRingDriver`` (default): ``mfrec_ring_*`` of the C ABI -- ONE persistent launch per rank for
  any number of epochs; a finished column block is written straight into the next rank's copy of
  Q through cudaIpc-mapped peer memory (NVLink) and its counter released at system scope.  No
  kernel boundary, staging copy or collective between steps; NCCL only carries the 64-byte memory
  handles, the per-epoch scalar all-reduce and the final gather of the slabs.
* ``Ring`` + ``GpuBackend`` (``exchange="nccl"``): one launch per slab and a batched
  ncclSend / ncclRecv of the whole slab between steps -- the round-1 transport, kept as the
  baseline the peer ring is measured against, and as the backend-neutral logic the gloo tests
  drive on CPU.

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



def layout_signature(ratings):
    """What every rank of a ring must agree on: grid shape and the item partition."""
    import zlib
    _, ip = ratings.perms()
    return (ratings.B, ratings.W, ratings.G, ratings.max_cb_items, zlib.crc32(ip.tobytes()),
            tuple(ratings.slab_items(c) for c in range(ratings.G)))


def check_layout_agreement(dist, ratings):
    """Ranks exchange Q rows by PACKED item position, so they must have computed the same item
    partition (same B, W, slab bounds, relabelling).  mfrec_ratings_pack derives all of it from
    the global item degrees it is given; this raises instead of exchanging misaligned rows if a
    caller passed rank-local degrees or different row_blocks / workers."""
    sig = layout_signature(ratings)
    sigs = [None] * dist.get_world_size()
    dist.all_gather_object(sigs, sig)
    if any(x != sigs[0] for x in sigs):
        raise RuntimeError("DSGD ranks disagree on the item layout (B, W, G, widest block, crc32(item_perm), slab bounds): %r"
                           % (sigs,))
    return sig


class PeerRingDriver(object):
    """Epoch driver over ``mfrec_ring_*``: persistent launches, peer-memory hand-over."""

    def __init__(self, torch, dist, native, ctx, ratings, model, kernel, hp, rank, world):
        self.torch, self.dist, self.native, self.ctx = torch, dist, native, ctx
        self.R, self.M, self.kernel, self.hp = ratings, model, kernel, hp
        self.rank, self.world = rank, world
        if world > 1:
            check_layout_agreement(dist, ratings)
        self.ring = native.PeerRing(ratings, model, rank, world, ctx)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, self.ring.export_handle())
            self.ring.connect(handles)
            dist.barrier()          # every rank has mapped its neighbour before anyone launches
        else:
            self.ring.connect_local(self.ring)
        self.stream = torch.cuda.ExternalStream(ctx.stream)

    def epochs(self, n, se_tensor):
        """n epochs in one launch; se_tensor: float64 CUDA tensor with >= n elements that receives
        this rank's sums of squared errors."""
        self.ring.epochs(self.kernel, self.hp["lr"], self.hp["K_users"], self.hp["K_items"], self.hp["K_bias"],
                         n, se_tensor.data_ptr())

    def finish(self):
        """Wait for the launches, copy the item side back into the model and gather the slabs: every
        rank ends with the complete item factors, like Ring.gather_items."""
        torch, dist = self.torch, self.dist
        self.ring.wait()
        if self.world > 1:
            dist.barrier()          # nobody reads its block while a neighbour may still push into it
        self.ring.sync_model()
        if self.world == 1:
            return
        (q_ptr, ib_ptr, _p, _ub), (ni, _nu, kpad) = self.M.device_ptrs()
        dev = torch.device("cuda", torch.cuda.current_device())
        Q = torch.as_tensor(_DevArray(q_ptr, (ni, kpad), "<f4"), device=dev)
        ib = torch.as_tensor(_DevArray(ib_ptr, (ni,), "<f4"), device=dev)
        for c in range(self.world):
            a, b = self.R.slab_items(c)
            for t in (Q[a:b], ib[a:b]):
                buf = t.clone()     # pool memory is not registered with NCCL: broadcast through a torch buffer
                dist.broadcast(buf, src=c)
                t.copy_(buf)
        torch.cuda.synchronize()

def user_slices(torch, users, nu, world):
    """Contiguous user-id slices with ~equal rating counts (every rank computes the same bounds)."""
    deg = torch.bincount(users.long(), minlength=nu)
    cum = torch.cumsum(deg, 0)
    total = int(cum[-1].item())
    targets = torch.tensor([total * w // world for w in range(1, world)], device=users.device, dtype=cum.dtype)
    cuts = (torch.searchsorted(cum, targets) + 1).tolist() if world > 1 else []
    return [0] + [min(int(c), nu) for c in cuts] + [nu]


def _pack_slice(native, args, ctx, world, idx_d, r_d, nu_r, ni_tot, deg_np, k):
    """This rank's slice of the ratings in the ring layout (every rank arrives at the same item layout:
    it is derived from the global item degrees)."""
    # (the slab-at-a-time NCCL transport cannot merge item copies at the end of an epoch: no splitting)
    split = native.SPLIT_OFF if (args.exchange == "nccl" or args.no_split) else native.SPLIT_AUTO
    return native.Ratings(None, None, ni_tot, nu_r, ctx=ctx, device_ptrs=(idx_d.data_ptr(), r_d.data_ptr()),
                          nnz=int(idx_d.shape[0]), ratings_are_f32=True, k_hint=k, n_slabs=world,
                          row_blocks=args.row_blocks, workers=args.workers, item_degree=deg_np, split=split)


def _run_ring(torch, dist, native, args, ctx, rank, world, dev, idx_d, r_d, nu_r, ni_tot, nnz_total, deg_np,
              u0, v0, k, hp, ClockSampler, local, R=None):
    """Pack this rank's slice (unless the caller did), run warm-up + timed epochs; returns a dict of
    measurements plus the objects needed for the end-to-end arm."""
    if R is None:
        R = _pack_slice(native, args, ctx, world, idx_d, r_d, nu_r, ni_tot, deg_np, k)
    M = native.Model(k, ni_tot, nu_r, u0, v0, None, None, layout=R, ctx=ctx)
    n_ep = args.warmup + args.steps
    se = torch.zeros(n_ep, device=dev, dtype=torch.float64)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    if args.exchange == "peer":
        drv = PeerRingDriver(torch, dist, native, ctx, R, M, native.KERNEL_LINEAR, hp, rank, world)
        stream = drv.stream
        with ClockSampler(local) as clocks:
            drv.epochs(args.warmup, se)
            drv.ring.wait()
            torch.cuda.synchronize()
            dist.barrier()
            launches1 = ctx.launch_count
            clocks.begin()
            ev0.record(stream)
            drv.epochs(args.steps, se[args.warmup:])        # ONE launch for all timed epochs
            ev1.record(stream)
            drv.ring.wait()
            torch.cuda.synchronize()
            dist.barrier()
            clocks.end()
        drv.finish()
        exchange = ("peer memory: one persistent launch per rank for all %d timed epochs, column blocks pushed "
                    "into the next rank's HBM over NVLink (cudaIpc), counters at system scope" % args.steps)
    else:
        check_layout_agreement(dist, R)
        be = GpuBackend(torch, native, ctx, R, M, native.KERNEL_LINEAR, hp)
        ring = Ring(be, rank, world, dist)
        stream = be.stream

        def one_epoch(e):
            with torch.cuda.stream(stream):
                se[e] = ring.epoch()

        with ClockSampler(local) as clocks:
            for e in range(args.warmup):
                one_epoch(e)
            ctx.sync()
            torch.cuda.synchronize()
            dist.barrier()
            launches1 = ctx.launch_count
            clocks.begin()
            ev0.record(stream)
            for e in range(args.steps):
                one_epoch(args.warmup + e)
            ev1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            dist.barrier()
            clocks.end()
        with torch.cuda.stream(stream):
            ring.gather_items()
        torch.cuda.synchronize()
        exchange = "nccl: one launch per slab, batched ncclSend/ncclRecv of the slab between steps (round-1 transport)"
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # device time, max over ranks
    dist.all_reduce(se)
    rmse = torch.sqrt(se / nnz_total).cpu().tolist()
    return dict(R=R, M=M, ms=float(ms.item()), rmse=rmse, clocks=clocks.summary(),
                launches_timed=int(ctx.launch_count - launches1), launches_warmup=int(launches1 - launches0),
                exchange=exchange)


def _e2e_ring(torch, dist, native, args, ctx, rank, world, dev, idx_d, r_d, nu_r, ni_tot, nnz_total, deg_np,
              u0, v0, k, hp):
    """End to end at N GPUs: every call takes this rank's slice from pinned HOST arrays, packs it,
    uploads the factors, builds and connects the ring, runs one epoch, gathers the slabs and reads
    the model back into host arrays.  Wall clock between barriers, max over ranks, median of calls."""
    import time
    import numpy as np
    idx_h = torch.empty(tuple(idx_d.shape), dtype=torch.int32, pin_memory=True)
    r_h = torch.empty(tuple(r_d.shape), dtype=torch.float32, pin_memory=True)
    idx_h.copy_(idx_d)
    r_h.copy_(r_d)
    torch.cuda.synchronize()
    idx_n, r_n = idx_h.numpy(), r_h.numpy()
    se = torch.zeros(1, device=dev, dtype=torch.float64)
    per_call, d2h = [], 0
    calls = max(args.e2e_steps, 3)
    for c in range(2 + calls):                     # two untimed calls: pool / IPC warm-up
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        R2 = native.Ratings(idx_n, r_n, ni_tot, nu_r, ctx=ctx, k_hint=k, n_slabs=world, row_blocks=args.row_blocks,
                            workers=args.workers, item_degree=deg_np,
                            split=native.SPLIT_OFF if (args.exchange == "nccl" or args.no_split) else native.SPLIT_AUTO)
        M2 = native.Model(k, ni_tot, nu_r, u0, v0, None, None, layout=R2, ctx=ctx)
        if args.exchange == "peer":
            drv = PeerRingDriver(torch, dist, native, ctx, R2, M2, native.KERNEL_LINEAR, hp, rank, world)
            drv.epochs(1, se)
            drv.finish()
        else:
            be = GpuBackend(torch, native, ctx, R2, M2, native.KERNEL_LINEAR, hp)
            ring = Ring(be, rank, world, dist)
            with torch.cuda.stream(be.stream):
                se[0] = ring.epoch()
                ring.gather_items()
        u1, v1, ib1, ub1 = M2.read()
        tot = se.clone()
        dist.all_reduce(tot)
        last = float(torch.sqrt(tot / nnz_total).item())
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if c >= 2:
            per_call.append(float(dt.item()) * 1e3)
        d2h = u1.nbytes + v1.nbytes + ib1.nbytes + ub1.nbytes + 8
        if args.exchange == "peer":
            del drv
        del R2, M2
    med = float(np.median(per_call))
    return {"value": nnz_total / (med * 1e-3), "unit": "updates/s", "ms_per_step": med, "steps": calls,
            "ms_per_call": per_call, "statistic": "median of calls, max over ranks per call",
            "h2d_bytes_per_step": int(idx_n.nbytes + r_n.nbytes + u0.nbytes + v0.nbytes),
            "d2h_bytes_per_step": int(d2h), "epochs_per_call": 1, "last_rmse": last,
            "what": "per rank: pinned host slice -> pack -> factors up -> ring set-up -> 1 epoch -> slab gather -> factors down"}


def bench_multi_gpu(args, rank, world, local, nu, ni, nnz, k, hp, gpu_synth, ClockSampler,
                    bytes_per_update, measured_peaks):
    """bench.py's N > 1 arm.

    strong (the line's `value`; BASELINE configs[2] names the SAME 100M-rating problem at 1/2/4/8
            GPUs): the single-GPU data set (same seed) cut into N user slices of equal rating count;
            rank 0 also trains the whole set on one GPU and the line carries
            parity = {rmse_n, rmse_1, rel} of the last epoch
    weak   (secondary key `weak`): one Netflix-shaped user slice per GPU, N tiles of the item
            catalogue -- per-GPU work fixed, the regime DSGD is built for
    """
    import time

    import numpy as np
    import torch
    import torch.distributed as dist
    from mfrec_b200 import _native as native, synth

    dev = torch.device("cuda", local)
    ctx = native.Context(local)
    peak, peak_src = measured_peaks()
    bpu = bytes_per_update(k)
    t_setup = time.time()
    modes = ["strong", "weak"] if args.scaling == "both" else [args.scaling]
    lines = {}
    for mode in modes:
        torch.cuda.empty_cache()
        if mode == "strong":
            idx_all, r_all = gpu_synth(torch, dev, nu, ni, nnz, seed=0)          # the N = 1 data set
            bounds = user_slices(torch, idx_all[:, 0], nu, world)
            a, b = bounds[rank], bounds[rank + 1]
            mine = (idx_all[:, 0] >= a) & (idx_all[:, 0] < b)
            idx_d = idx_all[mine].contiguous()
            idx_d[:, 0] -= a
            r_d = r_all[mine].contiguous()
            deg = torch.bincount(idx_all[:, 1].long(), minlength=ni)
            nu_r, ni_tot, nnz_total = b - a, ni, nnz
            if rank != 0:
                del idx_all, r_all
            del mine
            u0, v0_all = synth.init_factors(nu, ni, k, seed=2)
            v0 = np.ascontiguousarray(v0_all[:, a:b])
        else:
            ni_tot, nu_r, nnz_total = ni * world, nu, nnz * world
            idx_d, r_d = gpu_synth(torch, dev, nu, ni_tot, nnz, seed=1000 + rank, item_tiles=world, item_seed=0)
            deg = torch.bincount(idx_d[:, 1].long(), minlength=ni_tot)
            dist.all_reduce(deg)
            u0, _ = synth.init_factors(1, ni_tot, k, seed=2)      # same item init on every rank
            _, v0 = synth.init_factors(nu, 1, k, seed=100 + rank)
        deg_np = deg.cpu().numpy()
        # Pack first and agree on the outcome before any rank enters the ring's collectives: a layout the
        # library cannot build (e.g. the weak-scaling tile of the Yahoo shape at 8 GPUs: 1.09 M items x
        # 1.8 M users per rank need a 71-bit sort key) fails the line's primary mode loudly, but only
        # marks a secondary mode unavailable.
        R_mode, pack_err = None, None
        try:
            R_mode = _pack_slice(native, args, ctx, world, idx_d, r_d, nu_r, ni_tot, deg_np, k)
        except native.MfrecError as exc:
            pack_err = str(exc)
        agreed = torch.tensor([0 if pack_err else 1], device=dev, dtype=torch.int32)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()) == 0:
            if mode == modes[0]:
                raise native.MfrecError(-5, pack_err or "another rank could not pack its slice")
            lines[mode] = {"scaling": mode, "unavailable": pack_err or "another rank could not pack its slice"}
            del R_mode, idx_d, r_d
            dist.barrier()
            continue
        m = _run_ring(torch, dist, native, args, ctx, rank, world, dev, idx_d, r_d, nu_r, ni_tot, nnz_total,
                      deg_np, u0, v0, k, hp, ClockSampler, local, R=R_mode)
        del R_mode
        R, M = m.pop("R"), m.pop("M")
        value = nnz_total * args.steps / (m["ms"] * 1e-3)
        achieved = value / world * bpu / 1e9
        a0, b0 = R.slab_items(0)
        kpad = M.device_ptrs()[1][2]
        line = {"value": value, "ms_per_step": m["ms"] / args.steps, "scaling": mode,
                "workload": ("%s-shaped %dx%d nnz=%d k=%d cut into %d user slices (BASELINE configs[2] at N GPUs)"
                             % (args.workload, nu, ni, nnz, k, world)) if mode == "strong" else
                            ("%s-shaped tile per GPU: %d users x %d items, nnz=%d, k=%d; total %d x %d, nnz=%d"
                             % (args.workload, nu, ni, nnz, k, nu * world, ni_tot, nnz_total)),
                "parallelism": "dsgd ring of %d; item slab = %d rows, %d B per step and rank" % (world, b0 - a0, (b0 - a0) * (4 * kpad + 4)),
                "exchange": m["exchange"],
                "schedule": "stratified B=%d W=%d slabs=%d sub-epochs/epoch=%d; hot-item copies: %d items trained as %d rows"
                            % (R.B, R.W, R.G, R.G * R.B, R.copies()[2], R.copies()[1] - ni_tot + R.copies()[2]),
                "roofline_frac_algorithmic_per_gpu": achieved / peak,
                "rmse_per_epoch": m["rmse"], "gpu_launches": m["launches_timed"],
                "launches_warmup": m["launches_warmup"], "clocks": m["clocks"]}
        if mode == "strong" and not args.no_e2e:
            line["e2e"] = _e2e_ring(torch, dist, native, args, ctx, rank, world, dev, idx_d, r_d, nu_r, ni_tot,
                                    nnz_total, deg_np, u0, v0, k, hp)
        if mode == "strong":
            # parity: the same data, seeds and epochs on ONE GPU (rank 0), compared epoch by epoch
            del R, M
            torch.cuda.empty_cache()
            par = None
            if rank == 0:
                R1 = native.Ratings(None, None, ni, nu, ctx=ctx, device_ptrs=(idx_all.data_ptr(), r_all.data_ptr()),
                                    nnz=nnz, ratings_are_f32=True, k_hint=k)
                M1 = native.Model(k, ni, nu, u0, v0_all, None, None, layout=R1, ctx=ctx)
                n_ep = args.warmup + args.steps
                se1 = torch.zeros(n_ep, device=dev, dtype=torch.float64)
                for e in range(n_ep):
                    M1.sgd_epoch(R1, native.KERNEL_LINEAR, hp["lr"], hp["K_users"], hp["K_items"], hp["K_bias"],
                                 sq_err_ptr=se1.data_ptr() + 8 * e)
                ctx.sync()
                rm1 = torch.sqrt(se1 / nnz).cpu().tolist()
                par = {"rmse_n": m["rmse"][-1], "rmse_1": rm1[-1], "rel": abs(m["rmse"][-1] - rm1[-1]) / rm1[-1],
                       "epochs": n_ep, "rmse_1_per_epoch": rm1,
                       "what": "running RMSE of the last of %d epochs: %d-GPU ring vs one GPU, same data / seeds / init" % (n_ep, world)}
                del R1, M1, idx_all, r_all
            line["parity"] = par
        lines[mode] = line
        del idx_d, r_d
        dist.barrier()
    main = lines[modes[0]]
    out = {"metric": "rating_updates_per_s", "value": main["value"], "unit": "updates/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
           "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": main["workload"], "parallelism": main["parallelism"], "exchange": main["exchange"],
                      "kernel": "train_linear_kernel", "schedule": main["schedule"],
                      "l2": "inputs exceed the 126 MB L2", "hyper": hp},
           "roofline": {"bound": "hbm", "achieved": main["value"] / world * bpu / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": main["roofline_frac_algorithmic_per_gpu"], "traffic": None, "peak_source": peak_src,
                        "kernel": "sgd_block_kernel", "per": "GPU", "algorithmic_bytes_per_update": bpu,
                        "note": "algorithmic bytes (SURVEY 8(d)); see the N=1 line for measured DRAM traffic and the issue bound"},
           "cpu_baseline": None, "e2e": main.get("e2e"),
           "gpu_launches": main["gpu_launches"], "launches_warmup": main["launches_warmup"],
           "clocks": main["clocks"], "rmse_per_epoch": main["rmse_per_epoch"],
           "parity": main.get("parity"), "setup_s": time.time() - t_setup}
    for mode in modes[1:]:
        out[mode] = lines[mode]
    dist.barrier()
    dist.destroy_process_group()
    return out if rank == 0 else None
