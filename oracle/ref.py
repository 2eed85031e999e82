"""The reference's own native kernels (TEST INFRASTRUCTURE ONLY).

``oracle/build_ref.sh`` cythonises ``mfrec/lib/{kmf_train,gd_estimator,als_implicit}.pyx`` unmodified
from the read-only reference checkout into ``oracle/_ref/``; this module just imports the
resulting extension modules.  On the GPU box only the prebuilt files exist.
"""
import importlib.util
import os
import sysconfig

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")


def _load(name):
    path = os.path.join(_DIR, name + _SUFFIX)
    if not os.path.exists(path):
        raise ImportError(
            "oracle/_ref/%s%s missing: run oracle/build_ref.sh where /root/reference exists"
            % (name, _SUFFIX))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available():
    return all(os.path.exists(os.path.join(_DIR, n + _SUFFIX))
               for n in ("kmf_train", "gd_estimator"))


_cache = {}


def kmf_train():
    """mfrec.lib.kmf_train: train_linear_kernel, train_logistic_kernel."""
    if "kmf" not in _cache:
        _cache["kmf"] = _load("kmf_train")
    return _cache["kmf"]


def als_implicit():
    """mfrec.lib.als_implicit: als_wrmf (None if it was not built)."""
    if "als" not in _cache:
        try:
            _cache["als"] = _load("als_implicit")
        except ImportError:
            _cache["als"] = None
    return _cache["als"]


def gd_estimator():
    """mfrec.lib.gd_estimator: estimator_loop_without_bias, _with_bias, _with_bias_dev, ..."""
    if "gd" not in _cache:
        _cache["gd"] = _load("gd_estimator")
    return _cache["gd"]
