import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _gpu_available():
    """A usable sm_100 device behind libmfrec_b200 (there is no CPU path to fall back to)."""
    try:
        from mfrec_b200 import _native
        _native.Context(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a box without a B200 skips the gpu-marked tests instead of failing them
    (with `-m gpu` on a GPU box nothing is skipped; a missing library there still fails loudly in
    the tests that check it)."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no sm_100 CUDA device: libmfrec_b200 has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
