#!/usr/bin/env python
"""SURVEY T7 on hardware: `world` processes x `world` GPUs train ONE shared data set, split by
users, through the peer-memory DSGD ring; compared with the single-GPU run of the same data.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/ring_check.py --out gpurun_out/ring_check_nN.json

Checks written to the JSON (tests/test_ring_gpu.py::test_ring_over_nvlink asserts them):
  ranks_agree_on_layout   every rank computed the same item partition (unequal user slices)
  rmse_ring vs rmse_single, probe_ring vs probe_single   within 0.5 % after `--epochs` epochs
  bit_identical_runs      two ring runs from the same seeds give identical factors
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--nu", type=int, default=40000)
    ap.add_argument("--ni", type=int, default=3000)
    ap.add_argument("--nnz", type=int, default=3_000_000)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=10)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from mfrec_b200 import _native as native, dsgd, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    hp = dict(lr=0.005, K_users=0.05, K_items=0.05, K_bias=0.007)
    d = synth.make_ratings(args.nu, args.ni, args.nnz, seed=0, shuffle_seed=3, probe_frac=0.1)
    idx, r = d["idx"], d["r"]
    nnz = idx.shape[0]
    # deliberately UNEQUAL user slices (the layout must still agree: it comes from global degrees)
    cuts = [0] + [int(args.nu * (w / world) ** 1.15) for w in range(1, world)] + [args.nu]
    a, b = cuts[rank], cuts[rank + 1]
    mine = (idx[:, 0] >= a) & (idx[:, 0] < b)
    idx_r = np.ascontiguousarray(idx[mine])
    idx_r[:, 0] -= a
    r_r = np.ascontiguousarray(r[mine])
    deg_i = np.bincount(idx[:, 1], minlength=args.ni).astype(np.int64)
    u0, v0 = synth.init_factors(args.nu, args.ni, args.k, seed=2)
    ctx = native.Context(local)

    def ring_run():
        R = native.Ratings(idx_r, r_r, args.ni, b - a, ctx=ctx, k_hint=args.k, n_slabs=world, item_degree=deg_i)
        M = native.Model(args.k, args.ni, b - a, u0, np.ascontiguousarray(v0[:, a:b]), None, None, layout=R, ctx=ctx)
        drv = dsgd.PeerRingDriver(torch, dist, native, ctx, R, M, native.KERNEL_LINEAR, hp, rank, world)
        se = torch.zeros(args.epochs, device=dev, dtype=torch.float64)
        half = args.epochs // 2
        drv.epochs(half, se)                                  # two launches: the counters carry over
        drv.epochs(args.epochs - half, se[half:])
        drv.finish()
        dist.all_reduce(se)
        u1, v1, ib1, ub1 = M.read()
        parts = [None] * world
        dist.all_gather_object(parts, (a, b, v1, ub1))
        v = np.zeros_like(v0)
        ub = np.zeros(args.nu)
        for (pa, pb, pv, pub) in parts:
            v[:, pa:pb], ub[pa:pb] = pv, pub
        sig = dsgd.layout_signature(R)
        return u1, v, ib1, ub, torch.sqrt(se / nnz).cpu().numpy(), sig

    run1 = ring_run()
    run2 = ring_run()
    sigs = [None] * world
    dist.all_gather_object(sigs, run1[5])
    items = [None] * world
    dist.all_gather_object(items, (float(np.abs(run1[0]).sum()), float(np.abs(run1[2]).sum())))
    if rank == 0:
        u1, v1, ib1, ub1, rm_ring, _ = run1
        same = all(np.array_equal(x, y) for x, y in zip(run1[:4], run2[:4]))
        # single GPU, same data / seeds / epochs
        R = native.Ratings(idx, r, args.ni, args.nu, ctx=ctx, k_hint=args.k)
        M = native.Model(args.k, args.ni, args.nu, u0, v0, None, None, layout=R, ctx=ctx)
        se = torch.zeros(args.epochs, device=dev, dtype=torch.float64)
        for e in range(args.epochs):
            M.sgd_epoch(R, native.KERNEL_LINEAR, hp["lr"], hp["K_users"], hp["K_items"], hp["K_bias"],
                        sq_err_ptr=se.data_ptr() + 8 * e)
        ctx.sync()
        rm_single = torch.sqrt(se / nnz).cpu().numpy()
        us, vs, ibs, ubs = M.read()
        probe_ring, _ = native.rmse_pairs("predict_linear", u1, v1, d["probe_idx"], d["probe_r"], 0.0, ib1, ub1, ctx=ctx)
        probe_single, _ = native.rmse_pairs("predict_linear", us, vs, d["probe_idx"], d["probe_r"], 0.0, ibs, ubs, ctx=ctx)
        res = {"world": world, "shape": [args.nu, args.ni, int(nnz), args.k], "epochs": args.epochs,
               "user_slices": cuts, "layout": "B=%d W=%d G=%d" % run1[5][:3],
               "ranks_agree_on_layout": all(s == sigs[0] for s in sigs),
               "all_ranks_hold_the_same_item_side": all(x == items[0] for x in items),
               "rmse_ring": float(rm_ring[-1]), "rmse_single": float(rm_single[-1]),
               "rmse_ring_per_epoch": rm_ring.tolist(), "rmse_single_per_epoch": rm_single.tolist(),
               "probe_ring": float(probe_ring[0]), "probe_single": float(probe_single[0]),
               "bit_identical_runs": bool(same)}
        print(json.dumps(res))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
