"""Restatement of the camera->voxel lift of snap/models/streetview_encoder.py. Test infrastructure.

All functions operate on ONE scene (no batch axis) unless stated; fp32 NumPy.
"""
from __future__ import annotations

import itertools
from typing import Optional, Tuple

import numpy as np

from . import geometry, grids

F = np.float32


def project_points_to_views(scene_t_view: geometry.Transform3D, camera: geometry.Camera, points: np.ndarray):
    """snap/models/streetview_encoder.py:42-59 vmapped over V views (out_axes=1).

    scene_t_view: R [V,3,3], t [V,3]; camera fields [V,2]; points [N,3].
    Returns p2d [N,V,2] in (row, col) order, vis [N,V], depth [N,V], rays [N,V,3].
    """
    pts_view = scene_t_view.inv.transform(np.broadcast_to(points, (scene_t_view.R.shape[0],) + points.shape))
    depth = pts_view[..., -1]
    dist = np.sqrt(np.sum(pts_view * pts_view, axis=-1, keepdims=True)).astype(F)
    rays = (pts_view / np.maximum(dist, F(1e-5))).astype(F)
    p2d, vis = camera.world2image(pts_view)
    p2d = p2d[..., ::-1]  # :59 xy -> ij
    return (np.swapaxes(p2d, 0, 1), np.swapaxes(vis, 0, 1), np.swapaxes(depth, 0, 1), np.swapaxes(rays, 0, 1))


def interpolate_views_all(f_images: np.ndarray, p2d: np.ndarray) -> np.ndarray:
    """:69-76.  f_images [V,H,W,D], p2d [N,V,2] -> [N,V,D] (bilinear, clamp, fp32 coordinates)."""
    V = f_images.shape[0]
    out = [grids.interpolate_nd(f_images[v], p2d[:, v])[0] for v in range(V)]
    return np.stack(out, axis=1)


def interpolate_views_selective(f_images: np.ndarray, p2d: np.ndarray, index: np.ndarray,
                                cast=lambda a: a) -> np.ndarray:
    """:80-105.  f_images [V,H,W,D], p2d [N,K,2], index [N,K] -> [N,K,D].
    `cast` models `point.astype(arrays.dtype)` (:88): identity for fp32 features."""
    size = np.asarray(f_images.shape[1:3], dtype=F)
    pt = cast(p2d.astype(F))
    # jnp.minimum(bf16 point, int32 size - 1) promotes the size to the feature dtype (exact below 256)
    pt = np.maximum(np.minimum(cast(pt - F(0.5)), cast((size - 1).astype(F))), 0).astype(F)
    lower = np.floor(pt).astype(np.int32)
    upper = lower + 1
    w_upper = cast((pt - lower).astype(F))
    w_lower = cast((F(1.0) - w_upper).astype(F))
    weights = [w_lower, w_upper]
    coords = [lower, upper]
    H, W = f_images.shape[1:3]
    total = None
    for i, j in itertools.product(range(2), repeat=2):
        w = cast((weights[i][..., 0] * weights[j][..., 1]).astype(F))
        # jax gathers clamp out-of-bounds indices (default mode for x[i] inside jit is 'fill'/clip ->
        # the upper tap can only be OOB where its weight is exactly 0)
        r = np.clip(coords[i][..., 0], 0, H - 1)
        c = np.clip(coords[j][..., 1], 0, W - 1)
        val = cast(w[..., None] * f_images[index, r, c].astype(F))
        total = val if total is None else cast(total + val)
    return total.astype(F)


def interpolate_depth_score(score_scales: np.ndarray, depth: np.ndarray, depth_min_max=(1.0, 32.0)) -> np.ndarray:
    """:109-124.  score_scales [..., S], depth [...] -> [...] (linear interp over log-depth bins)."""
    S = score_scales.shape[-1]
    mn, mx = F(depth_min_max[0]), F(depth_min_max[1])
    d = np.clip(depth.astype(F), mn, mx)
    t = (np.log((d / mn).astype(F)).astype(F) / np.log(F(mx / mn)).astype(F)).astype(F)
    index = (F(0.5) + (t * F(S - 1)).astype(F)).astype(F)
    # grids.interpolate_nd on a [S,1] array at index: shift by 0.5, clamp taps
    c = (index - F(0.5)).astype(F)
    lo = np.floor(c).astype(np.int64)
    w_hi = (c - lo).astype(F)
    w_lo = (F(1.0) - w_hi).astype(F)
    lo_c = np.clip(lo, 0, S - 1)
    hi_c = np.clip(lo + 1, 0, S - 1)
    a = np.take_along_axis(score_scales, lo_c[..., None], -1)[..., 0].astype(F)
    b = np.take_along_axis(score_scales, hi_c[..., None], -1)[..., 0].astype(F)
    return ((w_lo * a).astype(F) + (w_hi * b).astype(F)).astype(F)


def view_selection(points: np.ndarray, scene_t_view: geometry.Transform3D, vis: np.ndarray, num: int):
    """:127-138.  points [N,3], t [V,3], vis [N,V] -> indices [N,num] int32, min_dist [N]."""
    diff = (points[:, None, :] - scene_t_view.t[None, :, :]).astype(F)
    dist = np.sqrt(np.sum(diff * diff, axis=-1)).astype(F)
    dist = np.where(vis, dist, F(np.inf))
    min_dist = dist.min(-1)
    # lax.top_k(-dist): descending, ties -> lower index first == stable argsort of dist
    indices = np.argsort(dist, axis=-1, kind="stable")[:, :num].astype(np.int32)
    return indices, min_dist


def pool_multiview_features(feats: np.ndarray, valid: np.ndarray, scores: Optional[np.ndarray] = None,
                            add_minmax: bool = True, use_variance: bool = True,
                            rd=lambda a: a) -> Tuple[np.ndarray, np.ndarray]:
    """:141-178.  feats [N,V,D], valid [N,V], scores [N,V] -> stats [N,C], valid_any [N].
    `rd` models `.astype(feats.dtype)` of mean/var (:163-164)."""
    valid_any = valid.any(-1)
    valid_ = np.where(valid_any[..., None], valid, True)[..., None]  # double-where (:150-152)
    feats = feats.astype(F)
    if scores is None:
        cnt = valid_.sum(-2).astype(F)
        mean_ = (np.where(valid_, feats, 0).sum(-2) / cnt).astype(F)
        var_ = (np.where(valid_, (feats - mean_[..., None, :]) ** 2, 0).sum(-2) / cnt).astype(F)
    else:
        s = scores.astype(F)[..., None]
        # jax.nn.softmax(x, where, initial=0): shift by max(initial, max_where x)
        mx = np.maximum(np.where(valid_, s, -np.inf).max(-2, keepdims=True), F(0)).astype(F)
        e = np.where(valid_, np.exp((s - mx).astype(F)), F(0)).astype(F)
        weights = (e / e.sum(-2, keepdims=True)).astype(F)
        weights = np.where(valid_, weights, F(0))
        mean_ = np.sum(weights * feats, axis=-2).astype(F)
        var_ = np.sum(weights * (feats - mean_[..., None, :]) ** 2, axis=-2).astype(F)
        mean_, var_ = rd(mean_), rd(var_)
    stats = [mean_]
    if use_variance:
        stats.append(var_)
    if add_minmax:
        stats.append(np.where(valid_, feats, -np.inf).max(-2))
        stats.append(np.where(valid_, feats, np.inf).min(-2))
    if scores is not None:
        stats.append(np.where(valid_, scores.astype(F)[..., None], -np.inf).max(-2))
    stats = np.where(valid_any[..., None], np.concatenate(stats, -1), F(0)).astype(F)
    return stats, valid_any
