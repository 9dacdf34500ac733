"""Golden fixture for VerticalPooling (bev_mapper.py:40-88): the reference's OWN `VerticalPooling.__call__` executed under
the NumPy stand-in for jax on a stand-in `self` (config + a plain-function confidence head), for the modes 'max', 'sum',
'mean', 'softmax' and 'weighted'.  Run in the build container only:

    python tests/golden/make_golden_pooling.py     # writes tests/golden/vertical_pooling.npz
"""
import os
import sys
import types as pytypes

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")

import jaxshim  # noqa: E402

jaxshim.install(REF)
from snap.models import bev_mapper as bm, types as rtypes  # noqa: E402

F = np.float32
rng = np.random.default_rng(99)
cells, Z, C = 40, 7, 16
feats = rng.standard_normal((cells, Z, C)).astype(F)
valid = rng.random((cells, Z)) < 0.6
valid[0] = False          # a column nobody sees: zero plane, invalid
valid[1] = True
w = (rng.standard_normal((C, 1)) * 0.5).astype(F)
b = np.array([0.2], F)
out = dict(feats=feats, valid=valid, head_kernel=w, head_bias=b)
for mode in ("max", "sum", "mean", "softmax", "weighted"):
    fake = pytypes.SimpleNamespace(config=pytypes.SimpleNamespace(pooling=mode), pooling_ops=bm.VerticalPooling.pooling_ops,
                                   confidence_head=lambda f: (f @ w + b).astype(F))
    pred = bm.VerticalPooling.__call__(fake, rtypes.FeatureVolume(features=feats, valid=valid))
    out[f"{mode}_plane"], out[f"{mode}_valid"] = pred["plane"].features, pred["plane"].valid
    if "weights" in pred:
        out[f"{mode}_scores"], out[f"{mode}_weights"] = pred["scores"], pred["weights"]
np.savez_compressed(os.path.join(HERE, "vertical_pooling.npz"), **{k: np.asarray(v) for k, v in out.items()})
print({k: np.asarray(v).shape for k, v in out.items()})
