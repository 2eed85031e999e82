"""Python 3 restatement of the on-path slice of ``mfrec.recommendation``:
same class / method / function names, positional signatures, array conventions and in-place
semantics as the reference; the native loops behind them run on the GPU through
``libmfrec_b200.so``.  ``pymongo`` / ``sparsesvd`` / neo4j exporters are not part of the path."""
from mfrec_b200.recommendation.base import BaseRecommender, Error  # noqa: F401
from mfrec_b200.recommendation.mf import MFRecommender  # noqa: F401
from mfrec_b200.recommendation.kmf import KMFRecommender  # noqa: F401
from mfrec_b200.recommendation.gradient_descent import GDRecommender  # noqa: F401
from mfrec_b200.recommendation.wrmf import WRMFRecommender  # noqa: F401
