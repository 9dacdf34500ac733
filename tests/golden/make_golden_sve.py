"""Golden fixture for the WHOLE lift orchestration: the reference's OWN `StreetViewEncoder.__call__`
(streetview_encoder.py:217-287) executed under the NumPy stand-in for jax on a stand-in `self` (config + plain NumPy
proj / fusion MLPs, precomputed `image_feature_pyr` so that no Flax encoder is needed), for the all-views path
(V <= top_k) and the view-selection path (V > top_k, with max_view_distance).  Run in the build container only:

    python tests/golden/make_golden_sve.py     # writes tests/golden/sve_call_*.npz
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
from snap.utils import geometry  # noqa: E402
from snap.models import streetview_encoder as sve, types as rtypes  # noqa: E402

F = np.float32
D, S = 8, 6


class Cfg(dict):
    __getattr__ = dict.__getitem__


def mlp_fn(params, input_act):
    def f(x, train=False):
        x = np.asarray(x, F)
        i = 0
        while f"Dense_{i}" in params:
            if i > 0 or input_act:
                x = np.maximum(x, 0)
            x = (x @ params[f"Dense_{i}"]["kernel"] + params[f"Dense_{i}"]["bias"]).astype(F)
            i += 1
        return jaxshim.clamp_indexing(x)      # JAX clamps out-of-bounds gather indices (:98-101)
    return f


def rot_cam(yaw):
    fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0]); down = np.array([0.0, 0.0, -1.0]); right = np.cross(down, fwd)
    return np.stack([right, down, fwd], axis=1)


def case(tag, V, max_view_distance, seed, weighted=True, depth_mlp=False):
    rng = np.random.default_rng(seed)
    B, X, Y, Z, Hf, Wf, Cin = 1, 6, 5, 4, 12, 16, 10
    if not weighted:
        Cin = D      # do_weighted_fusion=False samples the encoder features themselves (no proj MLP, :227-230)
    proj = {"Dense_0": {"kernel": (rng.standard_normal((Cin, D + S)) * 0.4).astype(F), "bias": (rng.standard_normal(D + S) * 0.1).astype(F)}}
    fusion = {"Dense_0": {"kernel": (rng.standard_normal((2 * D + 1, 12)) * 0.3).astype(F), "bias": (rng.standard_normal(12) * 0.1).astype(F)},
              "Dense_1": {"kernel": (rng.standard_normal((12, D)) * 0.3).astype(F), "bias": (rng.standard_normal(D) * 0.1).astype(F)}}
    if not weighted:   # statistics = [mean | var] (no score_max row, :174-177)
        fusion["Dense_0"]["kernel"] = fusion["Dense_0"]["kernel"][: 2 * D]
    dmlp = {"Dense_0": {"kernel": (rng.standard_normal((D + 4, 9)) * 0.4).astype(F), "bias": (rng.standard_normal(9) * 0.1).astype(F)},
            "Dense_1": {"kernel": (rng.standard_normal((9, D)) * 0.4).astype(F), "bias": (rng.standard_normal(D) * 0.1).astype(F)}}
    cfg = Cfg(do_weighted_fusion=weighted, num_scale_bins=S, top_k_view_selection=4, feature_dim=D, depth_min_max=(1.0, 32.0),
              fusion_add_minmax=False, fusion_use_variance=True, depth_mlp=Cfg() if depth_mlp else None,
              max_view_distance=max_view_distance)
    fake = pytypes.SimpleNamespace(config=cfg, dtype=F, proj_mlp=mlp_fn(proj, True), fusion_mlp=mlp_fn(fusion, False),
                                   depth_mlp=mlp_fn(dmlp, False))
    f_img = rng.standard_normal((B, V, Hf, Wf, Cin)).astype(F)
    R = np.stack([rot_cam(np.pi / 2 + rng.uniform(-0.4, 0.4)) for _ in range(V)])[None].astype(F)
    t = np.stack([[0.6 + 0.25 * v, 0.1 + rng.uniform(-0.1, 0.1), 0.5] for v in range(V)])[None].astype(F)
    cam = geometry.Camera(wh=np.tile(F([Wf * 4, Hf * 4]), (B, V, 1)), f=np.tile(F([40, 40]), (B, V, 1)),
                          c=np.tile(F([Wf * 2, Hf * 2]), (B, V, 1)))
    xs, ys, zs = (np.arange(X) + 0.5) * 0.4, (np.arange(Y) + 0.5) * 0.4 + 0.5, (np.arange(Z) + 0.5) * 0.3
    xyz = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1)[None].astype(F)
    f_in = f_img if weighted else jaxshim.clamp_indexing(f_img)   # JAX clamps out-of-bounds gather indices (:98-101)
    data = {"image_feature_pyr": rtypes.FeatureImagePyramid(features=[f_in], strides=[np.array([[4.0, 4.0]], F)]),
            "camera": cam, "T_view2scene": geometry.Transform3D(R=R, t=t), "xyz_query": xyz}
    pred = sve.StreetViewEncoder.__call__(fake, data, False)
    vol = pred["feature_volume"]
    out = dict(f_img=f_img, R=R, t=t, wh=cam.wh, f=cam.f, c=cam.c, xyz=xyz, volume=vol.features, valid=vol.valid,
               scores_images=pred.get("scores_images", np.zeros(0, F)), max_view_distance=np.asarray(-1.0 if max_view_distance is None else max_view_distance))
    for k, v in {"proj": proj, "fusion": fusion, **({"depth": dmlp} if depth_mlp else {})}.items():
        for n, p in v.items():
            out[f"{k}_{n}_kernel"], out[f"{k}_{n}_bias"] = p["kernel"], p["bias"]
    np.savez_compressed(os.path.join(HERE, f"sve_call_{tag}.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(tag, "valid voxels", int(np.asarray(vol.valid).sum()), "of", np.asarray(vol.valid).size, np.asarray(vol.features).shape)


case("allviews", 3, None, 1)
case("select", 6, 1.6, 2)
# do_weighted_fusion=False (:262-267): plain mean / variance pooling, without and with the per-observation depth_mlp
case("plain_allviews", 3, None, 3, weighted=False)
case("plain_select", 6, 1.6, 4, weighted=False)
case("plain_depthmlp_allviews", 3, None, 5, weighted=False, depth_mlp=True)
case("plain_depthmlp_select", 6, 1.6, 6, weighted=False, depth_mlp=True)
