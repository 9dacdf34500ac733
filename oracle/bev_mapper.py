"""Restatement of snap/models/bev_mapper.py (VerticalPooling, BEVMapper forward). Test infrastructure."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import geometry, grids, image_encoder, layers
from . import streetview_encoder as sv

F = np.float32
_id = lambda a: a


def vertical_pooling_max(features: np.ndarray, valid: np.ndarray):
    """bev_mapper.py:56-88 with pooling='max': max over axis -2 where valid (double-where), zero if none."""
    valid_any = valid.any(-1)
    vaa = np.where(valid_any[..., None], valid, True)
    pooled = np.where(vaa[..., None], features, -np.inf).max(-2)
    pooled = np.where(valid_any[..., None], pooled, F(0)).astype(features.dtype)
    return pooled, valid_any


def log_sigmoid(x: np.ndarray) -> np.ndarray:
    """jax.nn.log_sigmoid = -softplus(-x), fp32."""
    x = x.astype(F)
    return (-(np.maximum(-x, F(0)) + np.log1p(np.exp(-np.abs(x)).astype(F)).astype(F))).astype(F)


def vertical_pooling(features: np.ndarray, valid: np.ndarray, pooling: str = "max", params: Optional[Dict] = None,
                     rd: Callable = _id) -> Dict:
    """bev_mapper.py:56-88, all pooling modes.  features [..., Z, C] (already in the feature dtype), valid [..., Z]
    -> {'plane': (features [..., C], valid [...]), ['scores', 'weights' [..., Z] fp32]}.
    `rd` (numpy -> numpy) models the feature dtype: nn.Dense materialises dot and dot + bias in it (:64),
    `features.astype(feature dtype)` (:73), and half-precision sum / mean accumulate in fp32 and cast back."""
    valid = valid.astype(bool)
    valid_any = valid.any(-1)
    vaa = np.where(valid_any[..., None], valid, True)  # double-where (:58-60)
    out: Dict = {}
    f32 = features.astype(F)
    if pooling in ("weighted", "softmax"):
        head = params["confidence_head"]
        scores = rd(rd(layers.dense(f32, head["kernel"], None)) + head["bias"].astype(F))[..., 0].astype(F)
        if pooling == "weighted":
            scores = log_sigmoid(scores)
        out["scores"] = scores
        # jax.nn.softmax(where=, initial=0): shift by max(0, max over where)
        mx = np.maximum(np.where(vaa, scores, -np.inf).max(-1, keepdims=True), F(0)).astype(F)
        e = np.where(vaa, np.exp((scores - mx).astype(F)), F(0)).astype(F)
        w = (e / e.sum(-1, keepdims=True)).astype(F)
        w = out["weights"] = np.where(valid, w, F(0)).astype(F)
        pooled = rd(np.sum(f32 * w[..., None], axis=-2).astype(F))
    elif pooling == "mlp":
        x = np.where(valid[..., None], f32, F(0))
        x = x.reshape(*x.shape[:-2], -1)
        pooled = layers.mlp(x, params["fusion_mlp"], rd=rd)
    elif pooling == "max":
        pooled = np.where(vaa[..., None], f32, -np.inf).max(-2)
    elif pooling == "sum":
        pooled = rd(np.where(vaa[..., None], f32, F(0)).sum(-2).astype(F))
    elif pooling == "mean":
        cnt = vaa.sum(-1).astype(F)[..., None]
        pooled = rd((np.where(vaa[..., None], f32, F(0)).sum(-2).astype(F) / cnt).astype(F))
    else:
        raise NotImplementedError(pooling)  # :53-54
    pooled = np.where(valid_any[..., None], pooled, F(0)).astype(F)
    out["plane"] = (pooled, valid_any)
    return out


def bev_confidence(plane: np.ndarray, valid: np.ndarray, params: Dict, rd: Callable = _id) -> np.ndarray:
    """bev_mapper.py:292-295: where(valid, log_sigmoid(Dense(1)(plane)), 0); params = confidence_head tree."""
    head = params["layers_0"]
    s = rd(rd(layers.dense(plane.astype(F), head["kernel"], None)) + head["bias"].astype(F))[..., 0]
    return np.where(valid, log_sigmoid(s), F(0)).astype(F)


def build_xyz_query(grid: grids.Grid2D, t_view2scene_t: np.ndarray, scene_z_offset: float = 4.0,
                    scene_z_height: float = 12.0, z_offset: Optional[float] = None):
    """bev_mapper.py:159-196 for one scene (train=False): xyz [X,Y,Z,3] fp32."""
    xy = grid.index_to_xyz(grid.grid_index())  # [X,Y,2]
    if z_offset is None:
        cam_h = np.median(t_view2scene_t[..., -1].astype(F), axis=-1).astype(F)  # :171
        z_offset = F(cam_h - F(scene_z_offset))
    z = (np.arange(0, scene_z_height, grid.cell_size).astype(F) + F(z_offset)).astype(F)
    z = (z + F(grid.cell_size / 2)).astype(F)  # :189-193 (left-to-right adds)
    X, Y = grid.extent
    xyz = np.empty((X, Y, len(z), 3), dtype=F)
    xyz[..., :2] = xy[:, :, None, :]
    xyz[..., 2] = z[None, None, :]
    return xyz, F(z_offset)


def np_rd(rd_t: Callable) -> Callable:
    """Lift a torch rounding hook to numpy arrays."""
    if rd_t is _id:
        return _id
    return lambda a: rd_t(torch.from_numpy(np.ascontiguousarray(a, dtype=F))).numpy()


def lift_scene(f_proj_images: np.ndarray, camera: geometry.Camera, t_view2scene: geometry.Transform3D,
               xyz: np.ndarray, fusion_params: Dict, feature_dim: int = 128, top_k: int = 4,
               depth_min_max=(1.0, 32.0), rd: Callable = _id, chunk: int = 1 << 16,
               max_view_distance: Optional[float] = None, debug: Optional[Dict] = None,
               add_minmax: bool = False, use_variance: bool = True, threads: int = 1,
               weighted: bool = True, depth_mlp_params: Optional[Dict] = None):
    """streetview_encoder.py:232-286 for one scene, chunked over voxels (`threads` > 1: chunks on a thread pool, used by
    the timed CPU baseline of bench.py so that the lift uses the host cores like the torch-CPU encoder does).

    f_proj_images [V,Hf,Wf,feature_dim+S] = proj_mlp output; camera ALREADY scaled by 1/stride (:224).
    Returns f_grid [X,Y,Z,D], valid [X,Y,Z], and per-voxel debug (vis [N,V], p2d [N,V,2]); with view selection
    (V > top_k) vis / p2d are the GATHERED [N,top_k] arrays and `debug` (if given) collects 'view_indices'.

    weighted=False is the `do_weighted_fusion=False` branch (:262-267): `f_proj_images` [V,Hf,Wf,feature_dim] are the
    encoder features themselves (no proj MLP, no scale logits), the views are pooled with plain mean / variance
    (`scores=None`, :153-155) and, if `depth_mlp_params` is given, every observation first receives the residual
    `depth_mlp([f, log10(clip(depth, 0.1, 100)), ray])` (:264-267; rays zeroed where the view does not see the point).
    """
    rdn = np_rd(rd)
    grid_shape = xyz.shape[:-1]
    pts_all = xyz.reshape(-1, 3)
    V = f_proj_images.shape[0]

    def one_chunk(s):
        pts = pts_all[s:s + chunk]
        p2d, vis, depth, rays = sv.project_points_to_views(t_view2scene, camera, pts)
        min_dist, idx = None, None
        if top_k and V > top_k:  # :241-249
            idx, min_dist = sv.view_selection(pts, t_view2scene, vis, top_k)
            p2d, vis, depth, rays = (np.take_along_axis(a, idx[..., None] if a.ndim == 3 else idx, 1)
                                     for a in (p2d, vis, depth, rays))
            f_proj = sv.interpolate_views_selective(f_proj_images, p2d, idx, cast=rdn)
        else:
            f_proj = sv.interpolate_views_all(f_proj_images, p2d)
        f_proj = rdn(f_proj)
        if weighted:  # :254-261
            feats, scales = f_proj[..., :feature_dim], f_proj[..., feature_dim:]
            scores = rdn(sv.interpolate_depth_score(scales, depth, depth_min_max))
        else:  # :262-267
            feats, scores = f_proj, None
            if depth_mlp_params is not None:
                log_depth = np.log10(np.clip(depth, F(0.1), F(100))).astype(F)
                rays_ = np.where(vis[..., None], rays, F(0))
                f_depth = rdn(np.concatenate([feats, log_depth[..., None], rays_], -1))
                feats = rdn(feats + layers.mlp(f_depth, depth_mlp_params, rd=rdn))
        stats, valid = sv.pool_multiview_features(feats, vis, scores, add_minmax, use_variance, rd=rdn)  # :268-274
        if max_view_distance is not None and min_dist is not None:  # :275-279
            valid = valid & (min_dist <= F(max_view_distance))
        f = layers.mlp(rdn(stats), fusion_params, rd=rdn)  # :281
        f = np.where(valid[..., None], f, F(0))  # :282
        return f.astype(F), valid, vis, p2d, idx, stats

    if threads > 1:   # enough chunks to keep every thread busy (results do not depend on the chunking)
        chunk = max(4096, min(chunk, -(-len(pts_all) // (2 * threads))))
    starts = list(range(0, len(pts_all), chunk))
    if threads > 1 and len(starts) > 1:   # chunks are independent; NumPy releases the GIL inside its kernels
        import concurrent.futures as cf
        with cf.ThreadPoolExecutor(max_workers=min(threads, len(starts))) as ex:
            results = list(ex.map(one_chunk, starts))   # order preserved: identical to the sequential result
    else:
        results = [one_chunk(s) for s in starts]
    outs, valids, vis_all, p2d_all = [], [], [], []
    for f, valid, vis, p2d, idx, stats in results:
        if debug is not None:
            if idx is not None:
                debug.setdefault("view_indices", []).append(idx)
            debug.setdefault("stats", []).append(stats)
        outs.append(f); valids.append(valid); vis_all.append(vis); p2d_all.append(p2d)
    f_grid = np.concatenate(outs).reshape(*grid_shape, -1)
    valid = np.concatenate(valids).reshape(grid_shape)
    return f_grid, valid, np.concatenate(vis_all), np.concatenate(p2d_all)


def matching_head(plane: np.ndarray, valid: np.ndarray, p: Dict, rd: Callable = _id, normalize: bool = True):
    """bev_mapper.py:284-291: Dense 128->matching_dim, L2-normalise (if normalize_matching_features), mask."""
    rdn = np_rd(rd)
    f = rdn(rdn(layers.dense(plane, p["kernel"], None)) + p["bias"].astype(F))
    if normalize:
        f = rdn(layers.normalize(f))
    return np.where(valid[..., None], f, F(0)).astype(F)


def bev_mapper_forward(data: Dict, params: Dict, grid: grids.Grid2D, rd: Callable = _id,
                       scene_z_offset: float = 4.0, scene_z_height: float = 12.0, top_k: int = 4,
                       return_volume: bool = False, threads: int = 1, precomputed: Optional[Dict] = None,
                       feature_dim: int = 128, weighted: bool = True) -> Dict:
    """bev_mapper.py:254-296 for a batch (inference, train=False).

    data: 'images' f32 [B,V,H,W,3]; 'camera' geometry.Camera with fields [B,V,2];
          'T_view2scene' Transform3D R [B,V,3,3], t [B,V,3]; optional 'rasters' {'rgb' [B,G,G,3]}.
    params: Flax tree under 'bev_mapper'  with numpy leaves.
    precomputed (tests): {'sv_features' [B,V,Hf,Wf,C], 'sv_stride' (si, sj), 'aerial' [B,G,G,C]} replaces the image
    encoders (data['image_feature_pyr'] of streetview_encoder.py:218 and the aerial encoder output).
    """
    rdn = np_rd(rd)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F))
    tt = lambda tree: {k: (tt(v) if isinstance(v, dict) else t(v)) for k, v in tree.items()}
    B, V = (precomputed["sv_features"] if precomputed is not None else data["images"]).shape[:2]
    svp = params["streetview_encoder"]
    planes, valids, pred = [], [], {"streetview": []}
    for b in range(B):
        if precomputed is not None:
            feats, f_img, stride = [], precomputed["sv_features"][b], precomputed["sv_stride"]
        else:
            feats, strides = image_encoder.image_encoder(t(data["images"][b]), tt(svp["image_encoder"]), False, rd)
            f_img = feats[-1].numpy()
            stride = strides[-1]
        cam = geometry.Camera(wh=data["camera"].wh[b], f=data["camera"].f[b], c=data["camera"].c[b])
        cam = cam.scale(np.asarray([1 / stride[1], 1 / stride[0]], dtype=F))  # :224 (i,j) -> (x,y)
        if weighted:
            f_proj = layers.mlp(f_img, svp["proj_mlp"], apply_input_activation=True, rd=rdn)  # :229
        else:   # do_weighted_fusion=False: the encoder features are sampled as they are (:227-230 skipped)
            f_proj = np.asarray(f_img, dtype=F)
        T = geometry.Transform3D(R=data["T_view2scene"].R[b], t=data["T_view2scene"].t[b])
        xyz, z_off = build_xyz_query(grid, T.t, scene_z_offset, scene_z_height)
        f_grid, valid, vis, p2d = lift_scene(f_proj, cam, T, xyz, svp["fusion_mlp"], feature_dim=feature_dim, top_k=top_k, rd=rd,
                                             threads=threads, weighted=weighted, depth_mlp_params=svp.get("depth_mlp"))
        plane, pvalid = vertical_pooling_max(f_grid, valid)
        item = {"f_proj_images": f_proj, "feature_plane": plane, "valid": pvalid, "vis": vis, "p2d": p2d,
                "z_offset": z_off, "pyramid": [f.numpy() for f in feats]}
        if return_volume:
            item["feature_volume"], item["volume_valid"] = f_grid, valid
        pred["streetview"].append(item)
        planes_b, valids_b = [plane], [pvalid]
        if precomputed is not None and "aerial" in precomputed:
            aplane = precomputed["aerial"][b]
            pred.setdefault("aerial", []).append(aplane)
            planes_b.append(aplane); valids_b.append(np.ones(aplane.shape[:-1], bool))
        elif "rasters" in data and "aerial_encoder" in params:
            afeats, _ = image_encoder.image_encoder(t(data["rasters"]["rgb"][b:b + 1]), tt(params["aerial_encoder"]), True, rd)
            aplane = afeats[-1].numpy()[0]
            pred.setdefault("aerial", []).append(aplane)
            planes_b.append(aplane); valids_b.append(np.ones(aplane.shape[:-1], bool))
        if len(planes_b) > 1:  # fuse_neural_maps :225-252 == VerticalPooling('max') over the modality axis
            fused, fvalid = vertical_pooling_max(np.stack(planes_b, -2), np.stack(valids_b, -1))
        else:
            fused, fvalid = planes_b[0], valids_b[0]
        planes.append(fused); valids.append(fvalid)
    feats_bev = np.stack(planes); valid_bev = np.stack(valids)
    pred["bev_features"] = {"features": feats_bev, "valid": valid_bev}
    pred["bev_matching"] = {"features": matching_head(feats_bev, valid_bev, params["matching_proj"], rd),
                            "valid": valid_bev}
    return pred
