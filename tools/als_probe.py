"""ALS-WRMF seconds per epoch at ML-20M shape (bench.py secondary_als) as a stand-alone probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, json
from mfrec_b200 import _native
dev=torch.device("cuda",0); torch.cuda.set_device(0)
ctx=_native.default_context(0)
print(json.dumps(bench.secondary_als(torch, dev, _native, ctx, 0, with_cpu=False)))
os._exit(0)
