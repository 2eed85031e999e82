"""Seeded synthetic ratings of the shapes BASELINE.json names.

There is no network for the real MovieLens / Netflix / Yahoo sets, so every test and
benchmark uses ratings drawn from a planted low-rank model with power-law marginals:

* user activity  ~ log-normal(sigma = 1.0)
* item popularity ~ shifted Zipf  1 / (rank + q)^s  with s = 1 and q chosen so the most
  popular item holds ~0.25 % of all ratings (the Netflix Prize figure: 232 k of 100 M).
  A pure Zipf(1) would put 9.6 % of the mass on one item, i.e. more ratings than there
  are users once duplicates are removed.
* pairs are unique (the reference stores ratings in a scipy ``lil_matrix`` where a
  duplicate overwrites, base.py:823-828)
* values r = clip(round(3.6 + b_u + b_i + p_u . q_i + N(0, 0.5)), 1, 5), rank-16 planted model

Seeds: structure 0, planted model 1, factor init 2, order shuffle 3 (numpy PCG64), so the
same arrays can be handed to the oracle and to the CUDA path.
"""
import numpy as np

SHAPES = {
    # name: (nbr_users, nbr_items, nnz, k)
    "ml100k": (943, 1682, 100_000, 20),
    "ml20m": (138_000, 27_000, 20_000_000, 64),
    "netflix": (480_000, 17_700, 100_000_000, 128),
    "yahoo": (1_800_000, 136_000, 700_000_000, 128),
}


def marginals(nu, ni, seed=0, item_shift=None, top_item_share=0.0025):
    """Sampling weights (user, item), each summing to 1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    wu = rng.lognormal(mean=0.0, sigma=1.0, size=nu)
    wu /= wu.sum()
    if item_shift is None:
        # solve 1 / (q * ln((ni + q) / q)) ~= top_item_share for q by bisection
        lo, hi = 1e-3, float(ni) * 10
        for _ in range(100):
            mid = 0.5 * (lo + hi)
            share = (1.0 / (1.0 + mid)) / np.sum(1.0 / (np.arange(1, ni + 1) + mid))
            if share > top_item_share:
                lo = mid
            else:
                hi = mid
        item_shift = 0.5 * (lo + hi)
    wi = 1.0 / (np.arange(1, ni + 1, dtype=np.float64) + item_shift)
    wi = wi[rng.permutation(ni)]
    wi /= wi.sum()
    import os
    if os.environ.get("MFREC_SYNTH_UNIFORM"):   # experiment: no popularity skew at all
        wu = np.full(nu, 1.0 / nu)
        wi = np.full(ni, 1.0 / ni)
    return wu, wi


def sample_pairs(nu, ni, nnz, seed=0, **kw):
    """Unique (user, item) pairs with the power-law marginals; sorted by (user, item),
    i.e. scipy's lil->coo order (base.py:284-286)."""
    if nnz > nu * ni:
        raise ValueError("nnz exceeds the matrix size")
    wu, wi = marginals(nu, ni, seed, **kw)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    cu, ci = np.cumsum(wu), np.cumsum(wi)
    keys = np.zeros(0, dtype=np.int64)
    while keys.shape[0] < nnz:
        need = nnz - keys.shape[0]
        m = int(need * 1.15) + 1024
        us = np.minimum(np.searchsorted(cu, rng.random(m)), nu - 1).astype(np.int64)
        it = np.minimum(np.searchsorted(ci, rng.random(m)), ni - 1).astype(np.int64)
        new = np.setdiff1d(np.unique(us * ni + it), keys, assume_unique=True)
        if new.shape[0] > need:
            new = rng.permutation(new)[:need]
        keys = np.union1d(keys, new)
    idx = np.empty((nnz, 2), dtype=np.int32)
    idx[:, 0] = keys // ni
    idx[:, 1] = keys % ni
    return idx


def planted_values(idx, nu, ni, seed=1, rank=16):
    rng = np.random.Generator(np.random.PCG64(seed))
    bu = rng.normal(0.0, 0.3, nu)
    bi = rng.normal(0.0, 0.3, ni)
    p = rng.normal(0.0, 0.35, (nu, rank))
    q = rng.normal(0.0, 0.35, (ni, rank))
    out = np.empty(idx.shape[0], dtype=np.float64)
    step = 1 << 22
    for a in range(0, idx.shape[0], step):
        us, it = idx[a:a + step, 0], idx[a:a + step, 1]
        val = 3.6 + bu[us] + bi[it] + np.einsum("nk,nk->n", p[us], q[it])
        val += rng.normal(0.0, 0.5, us.shape[0])
        out[a:a + step] = np.clip(np.rint(val), 1.0, 5.0)
    return out


def init_factors(nu, ni, k, seed=2, std=0.1):
    """svd_u [k, ni] (items), svd_v [k, nu] (users) ~ N(0, std): mf.py:124-133."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.normal(0.0, std, (k, ni))
    v = rng.normal(0.0, std, (k, nu))
    return np.ascontiguousarray(u), np.ascontiguousarray(v)


def make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3, probe_frac=0.0):
    """Returns dict(idx, r[, probe_idx, probe_r]); order shuffled once like
    BaseRecommender.get_ratings(randomize_order=True) (base.py:1126-1131)."""
    idx = sample_pairs(nu, ni, nnz, seed)
    r = planted_values(idx, nu, ni, seed + 1)
    out = {}
    if probe_frac > 0:
        rng = np.random.Generator(np.random.PCG64(seed + 2000))
        probe = rng.random(nnz) < probe_frac
        out["probe_idx"] = np.ascontiguousarray(idx[probe])
        out["probe_r"] = np.ascontiguousarray(r[probe])
        idx, r = idx[~probe], r[~probe]
    if shuffle_seed is not None:
        order = np.random.Generator(np.random.PCG64(shuffle_seed)).permutation(idx.shape[0])
        idx, r = idx[order], r[order]
    out["idx"] = np.ascontiguousarray(idx, dtype=np.int32)
    out["r"] = np.ascontiguousarray(r, dtype=np.float64)
    return out
