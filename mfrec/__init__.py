"""``mfrec`` -- the reference's own import paths, served by the B200 implementation.

The reference binds its native loops as ``mfrec.lib.kmf_train`` / ``mfrec.lib.gd_estimator`` /
``mfrec.lib.als_implicit`` (setup.py:42-46, lib/setup.py:7-10) and its users import
``mfrec.recommendation.*`` (kmf.py:15-18, gradient_descent.py:20-25, mf.py:19-23).  This package
makes exactly those module paths resolve to ``mfrec_b200``'s modules -- the SAME module objects,
not copies, so module state such as ``kmf_train.last_rmse`` is shared:

    from mfrec.lib.kmf_train import train_linear_kernel          # kmf_train.pyx:195
    from mfrec.lib.gd_estimator import estimator_loop_without_bias
    from mfrec.recommendation.kmf import KMFRecommender
    from mfrec.recommendation.metrics import test_predict_rating

``mfrec.lib.machinelearning.kmf_train`` -- the path kmf.py:18 of the reference asks for, which
does not exist in the reference tree -- is provided as well, so that file's import line works
unchanged.  Nothing here computes; without ``libmfrec_b200.so`` every call raises ImportError.
"""
import importlib
import sys
import types

_ALIASES = (
    "lib", "lib.kmf_train", "lib.gd_estimator", "lib.als_implicit", "lib.datasets",
    "recommendation", "recommendation.base", "recommendation.mf", "recommendation.kmf",
    "recommendation.gradient_descent", "recommendation.wrmf", "recommendation.metrics",
)

for _name in _ALIASES:
    _mod = importlib.import_module("mfrec_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod

# kmf.py:18 of the reference: `from mfrec.lib.machinelearning.kmf_train import ...`
_ml = types.ModuleType(__name__ + ".lib.machinelearning")
_ml.__doc__ = "Alias package for the import path used by the reference's kmf.py:18."
_ml.__path__ = []
_ml.kmf_train = sys.modules[__name__ + ".lib.kmf_train"]
sys.modules[_ml.__name__] = _ml
sys.modules[_ml.__name__ + ".kmf_train"] = _ml.kmf_train
del _name, _mod, _ml
