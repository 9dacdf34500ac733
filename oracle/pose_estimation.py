"""Restatement of snap/models/pose_estimation.py (pose scoring, grid refinement, Kabsch, RANSAC sampling) and of the
matching part of snap/models/bev_localizer.py:137-218.  Test infrastructure (see oracle/__init__.py).

Random sampling (`jax.random.choice`, threefry) cannot be reproduced without JAX: `sample_transforms_ransac` takes
the sampled correspondence indices as an argument (or draws them from a NumPy generator) -- everything downstream of
the draw follows the reference line by line.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import geometry, grids

F = np.float32


def interpolate_score_maps(scores: np.ndarray, points: np.ndarray, valid: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """:50-62.  scores [N,H,W] (one score map per point), points [...,N,2] in cell units, valid [H,W]
    -> interpolated scores [...,N], valid [...,N].  = grids.interpolate_nd per point (grids.py:116-137): bilinear at
    (p - 0.5) with clamped taps summed in the order (0,0),(0,1),(1,0),(1,1); a point is valid if it lies inside the
    map and none of its four clamped taps is invalid (0/NaN mask)."""
    N, H, W = scores.shape
    size = np.asarray([H, W])
    p = points.astype(F)
    inb = np.all((p >= 0) & (p < size), -1)
    c = (p - F(0.5)).astype(F)
    lo = np.floor(c).astype(np.int64)
    w_hi = (c - lo).astype(F)
    w_lo = (F(1.0) - w_hi).astype(F)
    n_idx = np.broadcast_to(np.arange(N), p.shape[:-1])
    out = np.zeros(p.shape[:-1], dtype=F)
    ok = inb.copy()
    for ci in (0, 1):
        for cj in (0, 1):
            r = np.clip(lo[..., 0] + ci, 0, H - 1)
            q = np.clip(lo[..., 1] + cj, 0, W - 1)
            w = ((w_hi[..., 0] if ci else w_lo[..., 0]) * (w_hi[..., 1] if cj else w_lo[..., 1])).astype(F)
            out = (out + (w * scores[n_idx, r, q].astype(F)).astype(F)).astype(F)
            ok &= valid[r, q]
    return out, ok


def transform_points(angle: np.ndarray, t: np.ndarray, xy: np.ndarray) -> np.ndarray:
    """Transform2D.transform (geometry.py:138-140) for poses [...]: t + R p with the einsum's j order."""
    return geometry.Transform2D(angle=np.asarray(angle, F), t=np.asarray(t, F)).transform(xy.astype(F))


def pose_scoring_many(angle: np.ndarray, t: np.ndarray, scores_points_all: np.ndarray, i_xy_points: np.ndarray,
                      valid_points: np.ndarray, valid_j: np.ndarray, grid: grids.Grid2D,
                      mask_out_of_bounds: bool) -> np.ndarray:
    """:65-85 vmapped over poses (:206).  angle [P], t [P,2] -> score [P]."""
    j_uv = (transform_points(angle, t, i_xy_points) / F(grid.cell_size)).astype(F)  # [P,N,2]
    s, vj = interpolate_score_maps(scores_points_all, j_uv, valid_j)
    vp = np.broadcast_to(valid_points, s.shape)
    if mask_out_of_bounds:
        vp = vp & vj
    return np.sum(np.where(vp, s, F(0)), axis=-1, dtype=F)


def kabsch_algorithm_2d(i_p: np.ndarray, j_p: np.ndarray):
    """:103-123: least-squares rigid 2D transform i_t_j with i_p ~ R j_p + t.  Returns (angle, t, valid, rssd)."""
    i_p, j_p = i_p.astype(F), j_p.astype(F)
    mu_i, mu_j = i_p.mean(0, dtype=F), j_p.mean(0, dtype=F)
    a, b = (i_p - mu_i).astype(F), (j_p - mu_j).astype(F)
    cov = np.einsum("ji,jk->ik", a, b).astype(F)
    u, s, vh = np.linalg.svd(cov)
    sign = np.sign(np.linalg.det(u @ vh)).astype(F)
    u = (u * np.array([1, sign], F)).astype(F)
    s = (s * np.array([1, sign], F)).astype(F)
    valid = s[1] > F(1e-16) * s[0]
    error = np.sum(np.sum(a ** 2 + b ** 2, axis=1)) - 2 * np.sum(s)
    rssd = np.sqrt(np.clip(error, 0, None)).astype(F)
    R = (u @ vh).astype(F)
    t = (mu_i - R @ mu_j).astype(F)
    angle = np.arctan2(R[1, 0], R[0, 0]).astype(F)  # Transform2D.from_R (geometry.py:103-107)
    return angle, t, valid, rssd


def sample_transforms_ransac(indices: np.ndarray, i_xy_p: np.ndarray, num_poses: int, num_retries: int,
                             grid: grids.Grid2D):
    """:126-165 downstream of the draw.  indices int [num_poses*num_retries*2, 3] = unravelled (point, i, j) samples
    of the correspondence distribution.  Returns (angle [num_poses], t [num_poses, 2]) of j_t_i."""
    pool_shape = (num_poses, num_retries, 2, 2)
    i_pool = i_xy_p[indices[..., 0]].reshape(pool_shape).astype(F)
    j_pool = grid.index_to_xyz(indices[..., 1:]).reshape(pool_shape).astype(F)
    if num_retries > 1:
        d_i = np.linalg.norm(np.diff(i_pool, axis=-2).squeeze(-2), axis=-1).astype(F)
        d_j = np.linalg.norm(np.diff(j_pool, axis=-2).squeeze(-2), axis=-1).astype(F)
        ratio = np.maximum(d_i / np.clip(d_j, 1e-5, None), d_j / np.clip(d_i, 1e-5, None))
        sel = np.argmin(ratio, axis=-1)
        i_pool = i_pool[np.arange(num_poses), sel]
        j_pool = j_pool[np.arange(num_poses), sel]
    else:
        i_pool, j_pool = i_pool.squeeze(1), j_pool.squeeze(1)
    out = [kabsch_algorithm_2d(j_pool[k], i_pool[k]) for k in range(num_poses)]
    return np.stack([o[0] for o in out]).astype(F), np.stack([o[1] for o in out]).astype(F)


def draw_correspondences(rng: np.random.Generator, prob_points: np.ndarray, num: int) -> np.ndarray:
    """Stand-in for `jax.random.choice(rng, N*H*W, (num,), replace=True, p=prob)` + unravel_index (:139-146);
    NumPy generator instead of threefry (sampling is parity-unpinned)."""
    p = prob_points.reshape(-1).astype(np.float64)
    flat = rng.choice(p.size, size=num, replace=True, p=p / p.sum())
    return np.stack(np.unravel_index(flat, prob_points.shape), -1)


def refinement_offsets():
    """:177-188: the 41 x 41 x 41 grid of (degrees, x, y) offsets (jnp.mgrid with float steps, cast to fp32)."""
    delta_p, delta_r, range_p, range_r = 0.2, 0.25, 4, 5
    sp = slice(-range_p, range_p + delta_p, delta_p)
    sr = slice(-range_r, range_r + delta_r, delta_r)
    off = np.mgrid[sr, sp, sp].astype(F)
    return off.shape[1:], off.reshape(3, -1).T


def grid_refinement(angle0, t0, scores_points_all, i_xy_points, valid_points, valid_j, grid: grids.Grid2D,
                    mask_out_of_bounds: bool):
    """:168-203.  Returns (angle, t) of the refined pose and the score volume [41,41,41]."""
    shape, off = refinement_offsets()
    T_off = geometry.Transform2D(angle=np.deg2rad(off[:, 0]).astype(F), t=off[:, 1:].astype(F))
    T0 = geometry.Transform2D(angle=np.full(len(off), angle0, F), t=np.tile(np.asarray(t0, F), (len(off), 1)))
    Ts = T0 @ T_off
    scores = np.concatenate([
        pose_scoring_many(Ts.angle[s:s + 4096], Ts.t[s:s + 4096], scores_points_all, i_xy_points, valid_points,
                          valid_j, grid, mask_out_of_bounds) for s in range(0, len(off), 4096)])
    best = int(np.argmax(scores))
    return Ts.angle[best], Ts.t[best], scores.reshape(shape)


# ------------------------------------------------------------------------------------------------------------------
# snap/models/bev_localizer.py
# ------------------------------------------------------------------------------------------------------------------
def build_query_frustum_grid(cell_size: float, depth: float):
    """bev_localizer.py:36-54 (filter_points_in_fov=False): grid, grid_p_view [2], q_xy_p [X,Y,2]."""
    width = 3 * depth // 2
    extent = tuple(int(round(e / cell_size)) for e in (width, depth))
    grid = grids.Grid2D(extent, cell_size)
    grid_p_view = np.array([width / 2, 0.0], F)
    q_xy_p = (grid.index_to_xyz(grid.grid_index()) - grid_p_view).astype(F)
    return grid, grid_p_view, q_xy_p


def point_similarities(f_p_q: np.ndarray, valid_points: np.ndarray, map_features: np.ndarray,
                       temperature: Optional[float], clip_negative: bool = True, conf_p: Optional[np.ndarray] = None,
                       rd=lambda a: a):
    """bev_localizer.py:156-175.  f_p_q [N,D], map_features [H,W,D] -> sim_points, prob_points fp32 [N,H,W].
    `rd` models the feature dtype of the einsum output (:157)."""
    sim = rd(np.einsum("nd,ijd->nij", f_p_q.astype(F), map_features.astype(F)).astype(F))
    if clip_negative:
        sim = np.maximum(sim, 0)
    sim = sim.astype(F)
    if temperature is not None:
        sim = (sim * np.exp(F(temperature)).astype(F)).astype(F)
    m = sim.max(axis=(-1, -2), keepdims=True)
    e = np.exp((sim - m).astype(F)).astype(F)
    prob = (e / e.sum(axis=(-1, -2), keepdims=True, dtype=F)).astype(F)
    if conf_p is not None:  # masked_softmax over the valid points (layers.py:36-42: empty mask -> all valid)
        mask = valid_points if valid_points.any() else np.ones_like(valid_points, dtype=bool)
        c = np.where(mask, conf_p.astype(F), -np.inf)
        w = np.exp(c - c.max()).astype(F)
        w = (w / w.sum()).astype(F)[:, None, None]
        prob, sim = (prob * w).astype(F), (sim * w).astype(F)
    else:
        nv = F(max(int(valid_points.sum()), 1))
        sim, prob = (sim / nv).astype(F), (prob / nv).astype(F)
    return sim, prob


def sample_correspondences_inverse_cdf(prob_points: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """`jax.random.choice(p=prob.reshape(-1), replace=True)` (:139-146) = `searchsorted(cumsum(p), total * (1 - u))`
    (SURVEY Appendix A), restated in float64 as the nested inverse CDF the flat (point, row, column) order factorises
    into: uniforms [K,2] (point; cell inside the point's map) -> indices [K,3]."""
    N, H, W = prob_points.shape
    p = prob_points.astype(np.float64)
    row_mass = p.reshape(N, -1).sum(-1)
    cdf_n = np.cumsum(row_mass)
    n = np.minimum(np.searchsorted(cdf_n, cdf_n[-1] * (1.0 - uniforms[:, 0].astype(np.float64)), side="left"), N - 1)
    out = np.zeros((len(uniforms), 3), dtype=np.int64)
    for k, (nk, u) in enumerate(zip(n, uniforms[:, 1].astype(np.float64))):
        c = np.cumsum(p[nk].reshape(-1))
        flat = min(int(np.searchsorted(c, c[-1] * (1.0 - u), side="left")), H * W - 1)
        out[k] = (nk, flat // W, flat % W)
    return out


def transform2d_from_transform3d(R: np.ndarray, t: np.ndarray):
    """geometry.py:103-111: Transform2D.from_Transform3D -> (angle, t[:2])."""
    return np.arctan2(R[..., 1, 0], R[..., 0, 0]).astype(F), np.asarray(t, F)[..., :2]


def magnitude(angle: np.ndarray, t: np.ndarray):
    """geometry.py:132-136."""
    dr = (np.rad2deg(np.abs(angle)).astype(F) % F(360)).astype(F)
    dr = np.minimum(dr, F(360) - dr)
    return dr, np.linalg.norm(t, axis=-1).astype(F)


def loss_metrics(scores: np.ndarray, samples_angle: np.ndarray, samples_t: np.ndarray, best_angle, best_t,
                 gt_angle, gt_t, threshold_remove_accurate_poses=None):
    """bev_localizer.py:244-278 for ONE example: scores [P1], samples (index 0 = ground truth)."""
    S = geometry.Transform2D(angle=np.asarray(samples_angle, F), t=np.asarray(samples_t, F))
    G = geometry.Transform2D(angle=np.full(len(scores), gt_angle, F), t=np.tile(np.asarray(gt_t, F), (len(scores), 1)))
    rel = S.inv @ G
    dr_s, dt_s = magnitude(rel.angle, rel.t)
    sc = scores.astype(F).copy()
    if threshold_remove_accurate_poses is not None:
        dr_min, dt_min = threshold_remove_accurate_poses
        remove = (dr_s < dr_min) & (dt_s < dt_min)
        remove[0] = False
        sc = np.where(remove, -np.inf, sc).astype(F)
    m = sc.max()
    nll = -(sc[0] - m - np.log(np.sum(np.exp(sc - m), dtype=F)))
    Bt = geometry.Transform2D(angle=np.asarray([best_angle], F), t=np.asarray([best_t], F))
    Gt = geometry.Transform2D(angle=np.asarray([gt_angle], F), t=np.asarray([gt_t], F))
    relb = Bt.inv @ Gt
    dr, dt = magnitude(relb.angle, relb.t)
    metrics = {"loc/err_max_position": dt[0], "loc/err_max_rotation": dr[0],
               "loc/recall_top1": int(np.argmax(scores)) == 0}
    for dt_t, dr_t in [(0.5, 1), (1, 2), (2, 4)]:
        metrics[f"loc/recall_samples_{dt_t}m_{dr_t}"] = np.mean(((dr_s < dr_t) & (dt_s < dt_t))[1:])
    return F(nll), metrics, dr_s, dt_s
