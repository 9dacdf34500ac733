"""B200 implementation of `snap/models/bev_localizer.py`: the localizer the reference trains with (forward pass,
loss and metrics).  Map and query BEVs come from `BEVMapper`; the matching block (`:156-218`) runs on the kernels of
`csrc/localizer.cu` through `snap_b200.pose_estimation`.

Host-side geometry here (frustum grid, Transform3D -> Transform2D of the ground truth) is fp32 NumPy in the
reference's operation order; everything per-point / per-pose runs on the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import bev_mapper, configs, ops, pose_estimation, types

F = np.float32


def build_query_frustum_grid(cell_size: float, depth: float, filter_points_in_fov: bool = False,
                             hfov_deg: Optional[float] = None) -> Tuple[types.Grid2D, np.ndarray, np.ndarray]:
    """`bev_localizer.py:36-54`: gravity-aligned grid bounding the query frustum -> (grid, grid_p_view [2],
    q_xy_p [X,Y,2] or, when filtered to the field of view, [N,1,2])."""
    width = 3 * depth // 2                                              # :43
    extent = tuple(int(round(e / cell_size)) for e in (width, depth))   # Grid2D.from_extent_meters (grids.py:47-50)
    grid = types.Grid2D(extent, cell_size)
    grid_p_view = np.array([width / 2, 0.0], dtype=F)
    ii, jj = np.meshgrid(np.arange(extent[0]), np.arange(extent[1]), indexing="ij")
    idx = np.stack([ii, jj], -1)
    qgrid_xy_p = ((idx.astype(F) + F(0.5)) * F(cell_size)).astype(F)    # index_to_xyz (grids.py:62-63)
    q_xy_p = (qgrid_xy_p - grid_p_view).astype(F)
    if filter_points_in_fov:                                            # :50-53
        angle = np.arctan2(q_xy_p[..., 0], q_xy_p[..., 1]).astype(F)
        max_angle = hfov_deg / 2
        q_xy_p = q_xy_p[np.abs(angle) < np.deg2rad(max_angle).astype(F)][:, None]
    return grid, grid_p_view, q_xy_p


def transform2d_from_transform3d(T: types.Transform3D) -> np.ndarray:
    """`geometry.py:103-111` (from_Transform3D / from_R): rows (angle, tx, ty), fp32."""
    angle = np.arctan2(T.R[..., 1, 0], T.R[..., 0, 0]).astype(F)
    return np.concatenate([angle[..., None], np.asarray(T.t, dtype=F)[..., :2]], -1).astype(F)


class BEVLocalizer:
    """Mirror of `snap.models.bev_localizer.BEVLocalizer(config, scene_config, grid_map, semantic_map_classes, dtype)`.
    `scene_config` only needs `.streetview_hfov_deg` (`data/types.py:67`)."""

    default_config = staticmethod(configs.bev_localizer)

    def __init__(self, config=None, scene_config=None, grid_map: types.Grid2D = None, semantic_map_classes=None,
                 dtype=torch.bfloat16):
        self.config = c = config if config is not None else configs.bev_localizer()
        self.grid_map = grid_map
        hfov = getattr(scene_config, "streetview_hfov_deg", 72.0) if scene_config is not None else 72.0
        self.grid_query, self.qgrid_p_q, self.q_xy_p = build_query_frustum_grid(
            grid_map.cell_size, c.query_frustum_depth, c.filter_points_in_fov, hfov)        # :69-76
        if c.add_confidence_map:
            raise NotImplementedError("Map confidence is not yet supported.")               # :88-89
        if c.add_confidence_query:
            c.bev_mapper.add_confidence = True                                              # :90-91
        self.bev_mapper = bev_mapper.BEVMapper(c.bev_mapper, grid_map, semantic_map_classes, dtype)
        self.bev_mapper_query = None
        if c.bev_mapper_query is not None:
            self.bev_mapper_query = bev_mapper.BEVMapper(c.bev_mapper_query, grid_map, semantic_map_classes, dtype)
        if self.q_xy_p.shape[1] != 1:
            # the reference squeezes axis 2 of [B,N,1,2] (:153) and therefore only runs with the field-of-view filter
            raise ValueError("BEVLocalizer needs filter_points_in_fov=True (bev_localizer.py:153 squeezes the point grid)")
        self._q_xy_dev: Dict = {}

    def init_params(self, mapper_params: Dict, mapper_query_params: Optional[Dict] = None) -> Dict:
        p = {"bev_mapper": mapper_params}
        if self.bev_mapper_query is not None:
            p["bev_mapper_query"] = mapper_query_params
        if self.config.add_temperature:
            p["temperature"] = np.asarray(self.config.init_temperature, dtype=F)            # :107-109
        return p

    def recover_dense_feature_plane(self, plane_sparse: types.FeaturePlane) -> types.FeaturePlane:
        """`:111-129`: scatter the field-of-view points back onto the dense frustum grid (one example)."""
        q = self.q_xy_p[:, 0]
        idx = np.floor((q + self.qgrid_p_q[:2]) / F(self.grid_query.cell_size)).astype(np.int64)  # xyz_to_index
        ix, iy = torch.from_numpy(idx[:, 0]), torch.from_numpy(idx[:, 1])
        D = plane_sparse.features.shape[-1]
        dev = plane_sparse.features.device
        feats = torch.zeros((*self.grid_query.extent, D), dtype=plane_sparse.features.dtype, device=dev)
        valid = torch.zeros(self.grid_query.extent, dtype=plane_sparse.valid.dtype, device=dev)
        feats[ix.to(dev), iy.to(dev)] = plane_sparse.features.reshape(len(idx), D)
        valid[ix.to(dev), iy.to(dev)] = plane_sparse.valid.reshape(len(idx))
        return types.FeaturePlane(features=feats, valid=valid)

    def apply(self, variables: Dict, data: Dict, train: bool = False, debug: bool = False,
              rngs: Optional[Dict] = None) -> Dict:
        """`:131-218`.  `rngs['sampling']`: CUDA torch.Generator for the correspondence draws."""
        if train:
            raise NotImplementedError("training (backward kernels) is a 'next' row of SURVEY.md §8(f)")
        c = self.config
        params = variables["params"] if "params" in variables else variables
        pred: Dict = {}
        pred["map"] = self.bev_mapper.apply({"params": params["bev_mapper"]}, data["map"], train, debug)
        qm = self.bev_mapper_query or self.bev_mapper
        qp = params["bev_mapper_query"] if self.bev_mapper_query is not None else params["bev_mapper"]
        pred["query"] = qm.apply({"params": qp}, {**data["query"], "xy_bev": self.q_xy_p}, train, debug, is_query=True)
        plane_map, plane_q = pred["map"]["bev_matching"], pred["query"]["bev_matching"]
        pred.update(self.match(params, plane_q, plane_map, pred["query"].get("bev_confidence"),
                               data.get("T_query2map"), rngs))
        return pred

    __call__ = apply

    def match(self, params: Dict, plane_q: types.FeaturePlane, plane_map: types.FeaturePlane,
              conf_q: Optional[torch.Tensor] = None, T_query2map: Optional[types.Transform3D] = None,
              rngs: Optional[Dict] = None, uniforms: Optional[torch.Tensor] = None) -> Dict:
        """The matching block `:144-216` on given BEV planes (plane_q [B,N,1,D], plane_map [B,H,W,D])."""
        c = self.config
        B = plane_map.features.shape[0]
        dev = plane_map.features.device
        D = plane_q.features.shape[-1]
        key = str(dev)
        if key not in self._q_xy_dev:
            self._q_xy_dev[key] = torch.from_numpy(np.ascontiguousarray(self.q_xy_p[:, 0])).to(dev)   # :153
        q_xy_p = self._q_xy_dev[key]
        valid_points = plane_q.valid.reshape(B, -1)
        f_p_q = plane_q.features.reshape(B, -1, D)
        temperature = float(np.asarray(params["temperature"], dtype=F)) if c.add_temperature else None
        conf_p = conf_q.reshape(B, -1) if c.add_confidence_query else None
        maps = pose_estimation.point_similarities(f_p_q, valid_points, plane_map.features, temperature,
                                                  c.clip_negative_scores, conf_p)           # :156-172
        # sample poses (:175-182)
        K = c.num_pose_samples * c.num_pose_sampling_retries * 2
        if uniforms is None:
            gen = (rngs or {}).get("sampling")
            uniforms = torch.rand((B, K, 2), dtype=torch.float32, device=dev, generator=gen)
        indices = pose_estimation.sample_correspondences(maps, uniforms)
        m_t_q = pose_estimation.transforms_from_correspondences(indices, q_xy_p, c.num_pose_samples,
                                                                c.num_pose_sampling_retries, self.grid_map)
        start_idx = 0
        if T_query2map is not None:                                                         # :183-187
            gt = torch.from_numpy(transform2d_from_transform3d(T_query2map)).to(dev)
            m_t_q = torch.cat([gt[:, None], m_t_q], 1).contiguous()
            start_idx = 1
        out = {"map_t_query_samples": m_t_q, "correspondences": indices, "similarity_maps": maps}
        out["scores_poses"] = scores = pose_estimation.pose_scoring_many_batched(
            m_t_q, maps, q_xy_p, plane_map.valid, self.grid_map, c.mask_score_out_of_bounds)  # :190-198
        best_idx = torch.empty((B,), dtype=torch.int32, device=dev)
        best = torch.empty((B, 3), dtype=torch.float32, device=dev)
        ops.argmax_rows(scores, start_idx, best_idx, m_t_q, best)                           # :200-204
        out["best_index"], out["map_t_query"] = best_idx, best
        if c.do_grid_refinement:                                                            # :206-216
            out["map_t_query_ransac"] = best
            out["map_t_query"], out["scores_grid_refine"] = pose_estimation.grid_refinement_batched(
                best, maps, q_xy_p, plane_map.valid, self.grid_map, c.mask_score_out_of_bounds)
        return out

    def loss_metrics_function(self, pred: Dict, data: Dict, model_params: Optional[Dict] = None):
        """`BEVLocalizerModel.loss_metrics_function` (`:244-278`): per-example losses and metrics (device tensors).
        Requires the ground-truth pose to have been prepended to the samples (data['T_query2map'] given to apply)."""
        c = self.config
        scores, samples = pred["scores_poses"], pred["map_t_query_samples"]
        B, P1 = scores.shape
        dev = scores.device
        gt = torch.from_numpy(transform2d_from_transform3d(data["T_query2map"])).to(dev)
        out = torch.empty((B, 7), dtype=torch.float32, device=dev)
        ops.loc_nll(scores, samples, pred["map_t_query"].contiguous(), gt, c.threshold_remove_accurate_poses, out)
        nll, dr, dt = out[:, 0], out[:, 1], out[:, 2]
        losses = {"localization/nll": nll, "total": nll}
        metrics = {"loc/err_max_position": dt, "loc/err_max_rotation": dr, "loc/recall_top1": out[:, 3] > 0}
        for t in [0.5, 1, 2, 5]:
            metrics[f"loc/recall_max_{t}m"] = dt < t
            metrics[f"loc/recall_max_{t}°"] = dr < t
        if c.add_temperature and model_params is not None:
            metrics["loc/temperature"] = np.repeat(np.asarray(model_params["temperature"], dtype=F), B)
        for k, (dt_t, dr_t) in enumerate([(0.5, 1), (1, 2), (2, 4)]):
            metrics[f"loc/recall_samples_{dt_t}m_{dr_t}°"] = out[:, 4 + k]
        return losses, metrics
