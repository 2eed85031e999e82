"""The drop-in boundary without a GPU: libmfrec_b200.so loads, exports every entry point that
include/mfrec_b200.h declares (and nothing the Python binding expects is missing), and refuses
loudly to work without a CUDA device -- there is no CPU path in the product."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mfrec_b200.h")
LIB = os.path.join(ROOT, "mfrec_b200", "libmfrec_b200.so")


def declared_functions():
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mfrec_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("libmfrec_b200.so is not built: run __graft_entry__.build()")
    return ctypes.CDLL(LIB)


def test_header_declares_the_entry_points():
    names = declared_functions()
    for must in ("mfrec_train_kmf", "mfrec_train_funk", "mfrec_train_als_wrmf", "mfrec_predict_pairs",
                 "mfrec_rmse_pairs", "mfrec_topn", "mfrec_topn_sweep", "mfrec_bias_stats",
                 "mfrec_ratings_pack", "mfrec_sgd_epoch", "mfrec_funk_loop_dev", "mfrec_funk_subloop",
                 "mfrec_funk_predictor_subloop", "mfrec_train_funk_learned_bias", "mfrec_ctx_create"):
        assert must in names, must


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, "declared in include/mfrec_b200.h but not exported: %s" % missing


def test_python_binding_and_header_agree(lib):
    from mfrec_b200 import _native
    declared = set(declared_functions())
    assert set(_native.EXPORTS) <= declared, sorted(set(_native.EXPORTS) - declared)
    assert all(hasattr(lib, n) for n in _native.EXPORTS)


def test_abi_version_and_no_cpu_fallback(lib):
    from mfrec_b200 import _native
    assert lib.mfrec_abi_version() >= 1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present: the refusal cannot be observed")
    with pytest.raises(_native.MfrecError) as e:
        _native.Context(0)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)
    # the drop-in modules must fail the same way, never compute on the host
    import numpy as np
    from mfrec_b200.lib import kmf_train
    u, v = np.zeros((2, 3)), np.zeros((2, 4))
    idx = np.array([[0, 0], [1, 2]], dtype=np.int32)
    r = np.array([3.0, 4.0])
    with pytest.raises(_native.MfrecError):
        kmf_train.train_linear_kernel(1, 2, 0.1, 0.01, 0.0, 0.0, 0.05, 0.05, 0.007, 0.0, u, v, idx, r,
                                      np.zeros(3), np.zeros(4))
    assert not u.any() and not v.any()


def test_opts_struct_matches_the_header():
    """`mfrec_opts` as the Python binding lays it out == the C declaration (field order and names):
    a field added on one side only would shift every later field."""
    import ctypes as C
    import re
    from mfrec_b200 import _native
    with open(os.path.join(ROOT, "include", "mfrec_b200.h")) as f:
        text = f.read()
    body = text[text.index("typedef struct mfrec_opts {"):text.index("} mfrec_opts;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"\b(?:int32_t|uint64_t)\s+(\w+)\s*;", body)
    assert declared == [name for name, _ in _native.Opts._fields_]
    assert C.sizeof(_native.Opts) == 48          # 6 x int32, uint64, 3 x int32 (+ tail padding)


def test_storage_option_is_validated_on_the_host():
    from mfrec_b200.lib import _buffers
    old = _buffers.options["storage"]
    try:
        for name, code in (("f32", 0), ("f16", 1), ("bf16", 2)):
            _buffers.options["storage"] = name
            assert _buffers.native_opts()["storage"] == code
        _buffers.options["storage"] = "fp8"
        with pytest.raises(ValueError):
            _buffers.native_opts()
    finally:
        _buffers.options["storage"] = old
