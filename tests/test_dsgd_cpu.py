"""Multi-rank DSGD ring on CPU: ``mfrec_b200.dsgd.Ring`` over gloo (world_size 2 and 3) with the
CPU oracle as the per-block update.  Checks the schedule (every (rank, slab) block exactly once
per epoch, slabs handed on correctly) by comparing with a single-process replay of the same
blocks: ranks own disjoint users and, in any step, disjoint item slabs, so the distributed run
must be BIT-IDENTICAL to the replay."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mfrec_b200 import dsgd, synth
from oracle import cpu

NU, NI, NNZ, K, EPOCHS = 90, 60, 2500, 6, 3
HP = (0.01, 0.05, 0.06, 0.007)


def _problem(world):
    d = synth.make_ratings(NU, NI, NNZ, seed=5, shuffle_seed=6)
    idx, r = d["idx"], d["r"]
    user_rank = idx[:, 0] % world                       # user slices
    slab_of_item = (np.arange(NI) * world) // NI        # contiguous item slabs
    return idx, r, user_rank, slab_of_item


class OracleBackend(object):
    """Per-rank state: my users' ratings, full-size factor arrays (only my users / the slab in
    hand are meaningful), slabs exchanged as torch CPU tensors."""

    def __init__(self, rank, world):
        idx, r, user_rank, slab_of_item = _problem(world)
        mine = user_rank == rank
        self.idx, self.r = idx[mine], r[mine]
        self.slab_of_item = slab_of_item
        self.u, self.v = synth.init_factors(NU, NI, K, seed=2)
        self.ib, self.ub = np.zeros(NI), np.zeros(NU)
        self.world = world
        self.se = 0.0
        # slab buffers live in torch tensors (item-major so a slab is a contiguous row range)
        self.ut = torch.from_numpy(np.ascontiguousarray(self.u.T))   # [NI, K]
        self.ibt = torch.from_numpy(self.ib)
        self.bounds = [(int(np.searchsorted(slab_of_item, c)), int(np.searchsorted(slab_of_item, c + 1)))
                       for c in range(world)]

    def process_slab(self, c):
        m = self.slab_of_item[self.idx[:, 1]] == c
        if not m.any():
            return
        u = np.ascontiguousarray(self.ut.numpy().T)
        rm = cpu.kmf_train("linear", 1, K, *HP, u, self.v, np.ascontiguousarray(self.idx[m]),
                           np.ascontiguousarray(self.r[m]), self.ibt.numpy(), self.ub)
        self.ut.copy_(torch.from_numpy(np.ascontiguousarray(u.T)))
        self.se += float(rm[0]) ** 2 * int(m.sum())

    def slab_tensors(self, c):
        a, b = self.bounds[c]
        return [self.ut[a:b], self.ibt[a:b]]

    def sq_err(self):
        se, self.se = self.se, 0.0
        return se


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = OracleBackend(rank, world)
    ring = dsgd.Ring(be, rank, world, dist)
    ses = []
    for _ in range(EPOCHS):
        se = torch.tensor([ring.epoch()], dtype=torch.float64)
        dist.all_reduce(se)
        ses.append(float(se.item()))
    ring.gather_items()
    # gather the user side (each rank trained only its own users)
    mine = torch.zeros(NU, dtype=torch.bool)
    mine[np.arange(NU) % world == rank] = True
    v = torch.from_numpy(be.v.copy())
    v[:, ~mine] = 0
    ub = torch.from_numpy(be.ub.copy())
    ub[~mine] = 0
    dist.all_reduce(v)
    dist.all_reduce(ub)
    if rank == 0:
        np.savez(out, u=be.ut.numpy().T, v=v.numpy(), ib=be.ibt.numpy(), ub=ub.numpy(), se=np.array(ses))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _replay(world):
    idx, r, user_rank, slab_of_item = _problem(world)
    u, v = synth.init_factors(NU, NI, K, seed=2)
    ib, ub = np.zeros(NI), np.zeros(NU)
    ses = []
    for _ in range(EPOCHS):
        se = 0.0
        for step in range(world):
            for rank in range(world):
                c = dsgd.slab_at(rank, step, world)
                m = (user_rank == rank) & (slab_of_item[idx[:, 1]] == c)
                if m.any():
                    rm = cpu.kmf_train("linear", 1, K, *HP, u, v, np.ascontiguousarray(idx[m]),
                                       np.ascontiguousarray(r[m]), ib, ub)
                    se += float(rm[0]) ** 2 * int(m.sum())
        ses.append(se)
    return u, v, ib, ub, np.array(ses)


def test_schedule_is_a_latin_square():
    for world in (1, 2, 3, 4, 8):
        for step in range(world):
            slabs = [dsgd.slab_at(r, step, world) for r in range(world)]
            assert sorted(slabs) == list(range(world))          # no slab in two hands
        for r in range(world):
            assert sorted(dsgd.slab_at(r, s, world) for s in range(world)) == list(range(world))
            dst, src = dsgd.ring_peers(r, world)
            for step in range(world):
                # what I send after step t is what my destination works on in step t + 1
                assert dsgd.slab_at(dst, step + 1, world) == dsgd.slab_at(r, step, world)
                assert dsgd.slab_at(src, step, world) == dsgd.slab_at(r, step + 1, world)


@pytest.mark.parametrize("world", [2, 3])
def test_ring_over_gloo_matches_replay(world, tmp_path):
    out = str(tmp_path / "result.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    u, v, ib, ub, ses = _replay(world)
    assert np.array_equal(got["u"], u)
    assert np.array_equal(got["v"], v)
    assert np.array_equal(got["ib"], ib)
    assert np.array_equal(got["ub"], ub)
    np.testing.assert_allclose(got["se"], ses, rtol=1e-12)
