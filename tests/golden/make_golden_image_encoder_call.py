"""Golden fixture for the wrapper logic of ImageEncoder: the reference's OWN `ImageEncoder.__call__`
(image_encoder.py:119-144: pad_to_multiple, last unit of every stage, coarse-to-fine order, strides, crop to
ceil(input / stride)) under the NumPy stand-in for jax, with a stand-in encoder (average pooling pyramids) and an identity
decoder on a stand-in `self`.  Run in the build container only:

    python tests/golden/make_golden_image_encoder_call.py     # writes tests/golden/image_encoder_call.npz
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
from snap.models import image_encoder as ie  # noqa: E402

F = np.float32
rng = np.random.default_rng(31)


def pool(x, s):   # [B,H,W,C] -> [B,H/s,W/s,C]
    B, H, W, C = x.shape
    return x.reshape(B, H // s, s, W // s, s, C).mean((2, 4)).astype(F)


def fake_encoder(skip_root):
    base = 1 if skip_root else 4
    def enc(img, train=False):
        return {f"block{k + 1}": {"unit01": np.zeros((1,), F), "unit02": pool(img, base * 2 ** k)} for k in range(4)}
    return enc


out = {}
for tag, (hw, skip_root) in {"sv_30x44": ((30, 44), False), "sv_64x32": ((64, 32), False), "aerial_20x24": ((20, 24), True)}.items():
    img = rng.random((2, *hw, 3)).astype(F)
    fake = pytypes.SimpleNamespace(dtype=F, max_stride=(0 if skip_root else 2) + 3, level_names=["block4", "block3", "block2", "block1"],
                                   encoder=fake_encoder(skip_root), decoder=lambda feats, train=False: feats)
    pyr = ie.ImageEncoder.__call__(fake, img, False)
    out[f"{tag}_image"] = img
    for k, (f, s) in enumerate(zip(pyr.features, pyr.strides)):
        out[f"{tag}_feat{k}"], out[f"{tag}_stride{k}"] = f, s
    print(tag, [np.asarray(f).shape for f in pyr.features], [tuple(np.asarray(s)) for s in pyr.strides])
np.savez_compressed(os.path.join(HERE, "image_encoder_call.npz"), **{k: np.asarray(v) for k, v in out.items()})
