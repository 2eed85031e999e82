import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_problem():
    """300 users x 200 items, 6000 ratings, shuffled once (like get_ratings)."""
    from mfrec_b200 import synth
    nu, ni, nnz = 300, 200, 6000
    d = synth.make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3)
    return dict(nu=nu, ni=ni, nnz=nnz, idx=d["idx"], r=d["r"])


@pytest.fixture(scope="session")
def ml100k_problem():
    """BASELINE.json configs[0]: MovieLens-100K-shaped, k = 20, 90/10 train/probe split."""
    from mfrec_b200 import synth
    nu, ni, nnz, k = synth.SHAPES["ml100k"]
    d = synth.make_ratings(nu, ni, nnz, seed=0, shuffle_seed=3, probe_frac=0.1)
    d.update(nu=nu, ni=ni, k=k)
    return d
