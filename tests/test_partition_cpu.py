"""Host logic of the packer (no GPU): the balanced partition of users / items into blocks and groups
(mfrec_b200/csrc/partition.h) is a bijection onto contiguous id ranges, deterministic, and balances
power-law degrees to within a fraction of a percent."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("part") / "partition_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "mfrec_b200", "csrc"),
                           os.path.join(ROOT, "tools", "partition_check.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("n,nblocks,W,slabs", [(100000, 37, 8, 1), (480000, 148, 8, 1), (17700, 148, 8, 1),
                                              (40000, 24, 4, 4), (50, 3, 4, 1), (7, 1, 1, 1)])
def test_partition_is_a_balanced_bijection(checker, n, nblocks, W, slabs):
    out = subprocess.check_output([checker, str(n), str(nblocks), str(W), str(slabs)], text=True).split()
    assert out[0] == "ok", " ".join(out)
    gmax, bmax = float(out[1]), float(out[2])
    if n >= 40000 and n // (nblocks * W) >= 30:      # many light ids per group: near-perfect balance
        assert gmax < 1.01 and bmax < 1.005, out
