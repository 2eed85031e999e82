"""Would a finer hand-over shorten an epoch of sgd_block_kernel?  A timing model, no GPU.

The kernel's schedule: B CTAs x W warps; in sub-epoch s CTA rb holds column block
(rb + s) mod B, in phase p warp w works on column group (w + p) mod W of it.  Dependencies today:
  * a phase of warp w starts when warp w + 1 has finished its previous phase (shared-memory flag);
  * a sub-epoch of CTA rb starts when the CTA itself AND CTA rb + 1 have finished the previous one
    (CTA barrier + release / acquire counter per column block), + a fixed hand-over H.
"Group flow" (tried in round 2, dropped): no CTA barrier, one counter per column GROUP -- warp j starts
a sub-epoch when warp j + 1 of its own CTA and warp j + 1 of CTA rb + 1 have finished the previous one.

With bucket times ~ Poisson(45 ratings) x 0.23 us and H = 4 us the model gives 18.6 ms (today's
structure) vs 18.3 ms (group flow) vs 15.9 ms (no waiting at all): the ring of warps and the ring of CTAs
re-synchronise the schedule whatever the granularity of the counters, so the finer hand-over is worth
< 2 % -- and its implementation measured 34 ms per epoch (ten warps per SM spinning on counters instead
of one thread: `profiles/r02t_groupflow_experiment.txt`).  What the model says would pay is less
VARIANCE between the buckets of a phase (the packer balances row / column group totals, not buckets).

    python tools/handover_sim.py [B W ratings_per_bucket us_per_rating H_us]
"""
import sys

import numpy as np


def simulate(B=148, W=10, mean=45.0, us=0.23, H=4.0, seed=0):
    rng = np.random.default_rng(seed)
    S = B
    d = rng.poisson(mean, size=(B, S, W, W)).astype(float) * us   # [cta][sub-epoch][warp][phase]

    def phases(start, s):
        C = start
        for p in range(W):
            if p > 0:
                C = np.maximum(C, np.roll(C, -1, axis=1))
            C = C + d[:, s, :, p]
        return C

    end = np.zeros(B)
    for s in range(S):   # CTA barrier + one counter per column block
        start = np.maximum(end, np.roll(end, -1)) + H
        end = phases(np.tile(start[:, None], (1, W)), s).max(axis=1) + 1.0
    today = end.max()

    last = np.zeros((B, W))
    for s in range(S):   # one counter per column group, no CTA barrier
        nb = np.roll(last, -1, axis=0)
        start = np.maximum(last, np.maximum(np.roll(last, -1, axis=1), np.roll(nb, -1, axis=1))) + H
        last = phases(start, s) + 1.0
    group_flow = last.max()
    return today, group_flow, d.sum() / (B * W) + S * H


if __name__ == "__main__":
    a = [float(x) for x in sys.argv[1:]]
    args = dict(zip(("B", "W", "mean", "us", "H"), a))
    for k in ("B", "W"):
        if k in args:
            args[k] = int(args[k])
    t, g, ideal = simulate(**args)
    print("epoch, us: today's hand-over %.0f | per-column-group hand-over %.0f | no waiting %.0f" % (t, g, ideal))
