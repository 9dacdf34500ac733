"""B200 implementation of `snap/models/pose_estimation.py` (pose scoring, RANSAC sampling, grid refinement) and of
the point x map similarity block of `snap/models/bev_localizer.py:156-172`.

The reference materialises `sim_points` and `prob_points` as fp32 [B,N,H,W].  Here the similarities live in HBM once,
as the bf16 tensor the reference's einsum produces before its fp32 cast (`bev_localizer.py:157-160`), together with
the per-point factors (`exp(temperature)`, `1/num_valid` or the confidence soft-max) and the soft-max statistics;
`SimilarityMaps` carries them between the functions below.  Poses are float32 rows (angle [rad], tx, ty) =
`geometry.Transform2D`.  All batched functions take a leading batch axis (the reference vmaps them, `:206-226`).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops, types

F = np.float32


@dataclasses.dataclass
class SimilarityMaps:
    """`sim_points` / `prob_points` of `bev_localizer.py:156-172`, factored:
    sim_points[b,n] = point_scale[b,n]/valid * sim[b,n];  prob_points[b,n] = w[b,n] * softmax(scale * sim[b,n])."""
    sim: torch.Tensor            # bf16 [B,N,H*W] = (relu of) the rounded einsum
    scale: float                 # exp(temperature) (1 if add_temperature is off)
    point_scale: torch.Tensor    # f32 [B,N] = valid ? scale * w : 0
    row_cdf: torch.Tensor        # f32 [B,N] inclusive prefix of w (w = 1/num_valid or masked_softmax(conf))
    row_max: torch.Tensor        # f32 [B,N] max of scale * sim
    chunk_sum: torch.Tensor      # f32 [B,N,H] sum_j exp(scale * sim - row_max) per map row
    row_sum: torch.Tensor        # f32 [B,N]
    H: int = 0
    W: int = 0
    from_confidence: bool = False   # w came from the query confidences (their head needs a gradient the backward lacks)

    def sim_points(self) -> torch.Tensor:
        """The reference's fp32 `sim_points` [B,N,H,W] (debugging / tests; not used by the kernels)."""
        B, N = self.point_scale.shape
        w = self.row_cdf.clone()
        w[:, 1:] -= self.row_cdf[:, :-1]
        return (self.sim.float() * self.scale * w[..., None]).view(B, N, self.H, self.W)


def point_similarities(f_p_q: torch.Tensor, valid_points: torch.Tensor, map_features: torch.Tensor,
                       temperature: Optional[float] = None, clip_negative_scores: bool = True,
                       conf_p: Optional[torch.Tensor] = None) -> SimilarityMaps:
    """`bev_localizer.py:156-172`: f_p_q bf16 [B,N,D], valid_points u8 [B,N], map_features bf16 [B,H,W,D],
    conf_p f32 [B,N] (query confidences, `add_confidence_query`)."""
    B, N, D = f_p_q.shape
    _, H, W, _ = map_features.shape
    dev = f_p_q.device
    sim = torch.empty((B, N, H * W), dtype=torch.bfloat16, device=dev)
    fq = f_p_q.contiguous()
    fm = map_features.contiguous().view(B, H * W, D)
    for b in range(B):   # einsum('nd,ijd->nij') -> feature dtype -> relu (:157-159)
        ops.gemm(fq[b], fm[b], sim[b], relu=clip_negative_scores)
    scale = float(np.exp(F(temperature)).astype(F)) if temperature is not None else 1.0     # :161-162
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    maps = SimilarityMaps(sim=sim, scale=scale, point_scale=f32(B, N), row_cdf=f32(B, N), row_max=f32(B, N),
                          chunk_sum=f32(B, N, H), row_sum=f32(B, N), H=H, W=W, from_confidence=conf_p is not None)
    ops.loc_softmax_stats(sim, H, W, scale, maps.row_max, maps.chunk_sum, maps.row_sum)     # :163
    ops.loc_point_weights(valid_points.contiguous(), conf_p.contiguous() if conf_p is not None else None, scale,
                          maps.point_scale, maps.row_cdf)                                    # :165-172
    return maps


def sample_correspondences(maps: SimilarityMaps, uniforms: torch.Tensor) -> torch.Tensor:
    """`pose_estimation.py:139-146`: uniforms f32 [B,K,2] in [0,1) -> i32 [B,K,3] unravelled (point, i, j) draws
    from `prob_points` (inverse CDF in the reference's flat order; the random stream itself is jax's threefry in
    the reference and torch's Philox here)."""
    B, K, _ = uniforms.shape
    idx = torch.empty((B, K, 3), dtype=torch.int32, device=uniforms.device)
    ops.loc_sample(maps.sim, maps.row_max, maps.chunk_sum, maps.row_cdf, uniforms.contiguous(), maps.H, maps.W,
                   maps.scale, idx)
    return idx


def transforms_from_correspondences(indices: torch.Tensor, i_xy_p: torch.Tensor, num_poses: int, num_retries: int,
                                    grid: types.Grid2D) -> torch.Tensor:
    """`pose_estimation.py:147-165` downstream of the draw: indices i32 [B, num_poses*num_retries*2, 3] ->
    j_t_i f32 [B,num_poses,3] (most consistent retry per pose, then `kabsch_algorithm_2d` :103-123)."""
    B = indices.shape[0]
    poses = torch.empty((B, num_poses, 3), dtype=torch.float32, device=indices.device)
    ops.loc_ransac_poses(indices.contiguous(), i_xy_p.contiguous(), num_poses, num_retries, float(F(grid.cell_size)),
                         poses)
    return poses


def sample_transforms_ransac_batched(rng: torch.Generator, maps: SimilarityMaps, i_xy_p: torch.Tensor, num_poses: int,
                                     num_retries: int, grid: types.Grid2D) -> torch.Tensor:
    """`pose_estimation.py:126-165,220-222`.  `rng`: a CUDA torch.Generator (the 'sampling' rng stream)."""
    B = maps.point_scale.shape[0]
    dev = maps.sim.device
    u = torch.rand((B, num_poses * num_retries * 2, 2), dtype=torch.float32, device=dev, generator=rng)
    return transforms_from_correspondences(sample_correspondences(maps, u), i_xy_p, num_poses, num_retries, grid)


def pose_scoring_many_batched(j_t_i: torch.Tensor, maps: SimilarityMaps, i_xy_points: torch.Tensor,
                              valid_j: Optional[torch.Tensor], grid: types.Grid2D,
                              mask_out_of_bounds: bool) -> torch.Tensor:
    """`pose_estimation.py:65-85,206-209`: j_t_i f32 [B,P,3] -> scores f32 [B,P].  `valid_points` is folded into
    `maps.point_scale`; valid_j u8 [B,H,W] is only read when mask_out_of_bounds."""
    B, P, _ = j_t_i.shape
    scores = torch.empty((B, P), dtype=torch.float32, device=j_t_i.device)
    ops.loc_pose_scoring(maps.sim, maps.point_scale, i_xy_points.contiguous(),
                         valid_j.contiguous() if valid_j is not None else None, j_t_i.contiguous(), maps.H, maps.W,
                         float(F(grid.cell_size)), mask_out_of_bounds, scores)
    return scores


def refinement_offsets() -> Tuple[np.ndarray, np.ndarray]:
    """`pose_estimation.py:177-184`: the axes of jnp.mgrid[slice_r, slice_p, slice_p] (degrees; metres)."""
    delta_p, delta_r, range_p, range_r = 0.2, 0.25, 4, 5
    slice_p = slice(-range_p, range_p + delta_p, delta_p)
    slice_r = slice(-range_r, range_r + delta_r, delta_r)
    off = np.mgrid[slice_r, slice_p, slice_p].astype(F)       # the reference's 3-D mgrid (:181); its axes:
    return np.ascontiguousarray(off[0][:, 0, 0]), np.ascontiguousarray(off[1][0, :, 0])


_REFINE_AXES: dict = {}


def grid_refinement_batched(j_t_i_init: torch.Tensor, maps: SimilarityMaps, i_xy_points: torch.Tensor,
                            valid_j: Optional[torch.Tensor], grid: types.Grid2D,
                            mask_out_of_bounds: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """`pose_estimation.py:168-203,210-212`: j_t_i_init f32 [B,3] -> (refined f32 [B,3], scores f32 [B,nr,np,np])."""
    B = j_t_i_init.shape[0]
    dev = j_t_i_init.device
    rot, pos = refinement_offsets()
    if str(dev) not in _REFINE_AXES:    # uploaded once per device (no H2D copy inside a captured graph)
        _REFINE_AXES[str(dev)] = (torch.from_numpy(np.deg2rad(rot).astype(F)).to(dev), torch.from_numpy(pos).to(dev))
    rot_rad, pos_t = _REFINE_AXES[str(dev)]
    P = len(rot) * len(pos) * len(pos)
    poses = torch.empty((B, P, 3), dtype=torch.float32, device=dev)
    ops.loc_refine_poses(j_t_i_init.contiguous(), rot_rad, pos_t, pos_t, poses)
    scores = pose_scoring_many_batched(poses, maps, i_xy_points, valid_j, grid, mask_out_of_bounds)
    best = torch.empty((B,), dtype=torch.int32, device=dev)
    refined = torch.empty((B, 3), dtype=torch.float32, device=dev)
    ops.argmax_rows(scores, 0, best, poses, refined)
    return refined, scores.view(B, len(rot), len(pos), len(pos))
