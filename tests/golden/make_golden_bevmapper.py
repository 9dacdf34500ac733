"""Golden fixture for BEVMapper downstream of the image encoders: the reference's OWN `BEVMapper.__call__`,
`encode_streetview`, `encode_aerial`, `fuse_neural_maps` (bev_mapper.py:159-296), `StreetViewEncoder.__call__` and
`VerticalPooling.__call__`, executed under the NumPy stand-in for jax on stand-in `self` objects (configs + plain NumPy
dense layers; the street-view pyramid comes in as data['image_feature_pyr'], the aerial encoder is a lookup of a
precomputed feature map).  Pins: voxel grid from the median camera height, lift, vertical max, modality max, matching
head, confidence.  Run in the build container only:

    python tests/golden/make_golden_bevmapper.py     # writes tests/golden/bevmapper_call.npz
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
from snap.utils import geometry, grids  # noqa: E402
from snap.models import bev_mapper as bm, streetview_encoder as sve, types as rtypes  # noqa: E402
from make_golden_sve import Cfg, mlp_fn, rot_cam  # noqa: E402  (also regenerates the sve fixtures: deterministic)

F = np.float32
D, S, DM = 8, 6, 4
rng = np.random.default_rng(123)
B, V, G, Hf, Wf, Cin = 2, 3, 6, 12, 16, 10
cell = 0.4
proj = {"Dense_0": {"kernel": (rng.standard_normal((Cin, D + S)) * 0.4).astype(F), "bias": (rng.standard_normal(D + S) * 0.1).astype(F)}}
fusion = {"Dense_0": {"kernel": (rng.standard_normal((2 * D + 1, 12)) * 0.3).astype(F), "bias": (rng.standard_normal(12) * 0.1).astype(F)},
          "Dense_1": {"kernel": (rng.standard_normal((12, D)) * 0.3).astype(F), "bias": (rng.standard_normal(D) * 0.1).astype(F)}}
Wm, bmv = (rng.standard_normal((D, DM)) * 0.5).astype(F), (rng.standard_normal(DM) * 0.1).astype(F)
wc, bc = (rng.standard_normal((D, 1)) * 0.5).astype(F), np.array([0.1], F)
sv_cfg = Cfg(do_weighted_fusion=True, num_scale_bins=S, top_k_view_selection=4, feature_dim=D, depth_min_max=(1.0, 32.0),
             fusion_add_minmax=False, fusion_use_variance=True, depth_mlp=None, max_view_distance=None)
fake_sve = pytypes.SimpleNamespace(config=sv_cfg, dtype=F, proj_mlp=mlp_fn(proj, True), fusion_mlp=mlp_fn(fusion, False))
fake_vp = pytypes.SimpleNamespace(config=pytypes.SimpleNamespace(pooling="max"), pooling_ops=bm.VerticalPooling.pooling_ops)
aerial = rng.standard_normal((B, G, G, D)).astype(F)
cfg = Cfg(scene_z_offset=1.0, scene_z_offset_range=(-2, 2), scene_z_height=1.6, matching_dim=DM, normalize_matching_features=True,
          add_confidence=True, apply_modality_dropout=True)
fake = pytypes.SimpleNamespace(
    config=cfg, grid=grids.Grid2D((G, G), cell),
    streetview_encoder=lambda data, train: sve.StreetViewEncoder.__call__(fake_sve, data, train),
    vertical_pooling=lambda vol: bm.VerticalPooling.__call__(fake_vp, vol),
    modality_fusion=lambda vol: bm.VerticalPooling.__call__(fake_vp, vol),
    aerial_encoder=lambda rgb, train: rtypes.FeatureImagePyramid(features=[aerial], strides=[None]),
    semantic_encoder=None,
    matching_proj=lambda f: (f @ Wm + bmv).astype(F),
    confidence_head=lambda f: (f @ wc + bc).astype(F))
for name in ("encode_streetview", "encode_aerial", "fuse_neural_maps"):
    setattr(fake, name, (lambda fn: lambda *a, **k: fn(fake, *a, **k))(getattr(bm.BEVMapper, name)))
f_img = rng.standard_normal((B, V, Hf, Wf, Cin)).astype(F)
R = np.stack([[rot_cam(np.pi / 2 + rng.uniform(-0.4, 0.4)) for _ in range(V)] for _ in range(B)]).astype(F)
t = np.stack([[[0.6 + 0.5 * v, 0.1 + rng.uniform(-0.1, 0.1), 1.3 + 0.2 * v + 0.1 * b] for v in range(V)] for b in range(B)]).astype(F)
cam = geometry.Camera(wh=np.tile(F([Wf * 4, Hf * 4]), (B, V, 1)), f=np.tile(F([40, 40]), (B, V, 1)), c=np.tile(F([Wf * 2, Hf * 2]), (B, V, 1)))
data = {"image_feature_pyr": rtypes.FeatureImagePyramid(features=[f_img], strides=[np.tile(F([4.0, 4.0]), (B, 1))]),
        "camera": cam, "T_view2scene": geometry.Transform3D(R=R, t=t), "rasters": {"rgb": np.zeros((B, G, G, 3), F)}}
pred = bm.BEVMapper.__call__(fake, data, False, False, False)
out = dict(f_img=f_img, R=R, t=t, wh=cam.wh, f=cam.f, c=cam.c, aerial=aerial, cell=np.asarray(cell), Wm=Wm, bm=bmv, wc=wc, bc=bc,
           xyz=data["xyz_query"], sv_plane=pred["streetview"]["feature_plane"].features,
           sv_valid=pred["streetview"]["feature_plane"].valid, bev_features=pred["bev_features"].features,
           bev_valid=pred["bev_features"].valid, bev_matching=pred["bev_matching"].features, bev_confidence=pred["bev_confidence"])
for k, v in {"proj": proj, "fusion": fusion}.items():
    for n, p in v.items():
        out[f"{k}_{n}_kernel"], out[f"{k}_{n}_bias"] = p["kernel"], p["bias"]
np.savez_compressed(os.path.join(HERE, "bevmapper_call.npz"), **{k: np.asarray(v) for k, v in out.items()})
print({k: np.asarray(v).shape for k, v in out.items() if not k.startswith(("proj", "fusion"))})
print("street-view valid cells", int(np.asarray(out["sv_valid"]).sum()), "of", np.asarray(out["sv_valid"]).size)
