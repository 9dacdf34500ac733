"""Restatement of snap/models/pose_exhaustive_voting.py. Test infrastructure (NumPy + torch-CPU conv)."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as Fnn

from . import geometry, grids

F = np.float32


def get_grid_center_transform(grid: grids.Grid2D) -> geometry.Transform2D:
    """:31-34: corner_t_center."""
    center_offset = (np.asarray(grid.extent_meters) / 2).astype(F)
    return geometry.Transform2D(angle=np.asarray(0, dtype=F), t=center_offset)


def template_transforms(num_rotations: int, grid: grids.Grid2D) -> geometry.Transform2D:
    """:45-50: templates_t_grid = corner_t_center @ rotated_t_grid @ corner_t_center.inv (fp32)."""
    angles = np.linspace(0, np.pi * 2, num_rotations, endpoint=False).astype(F)
    rotated = geometry.Transform2D(angle=angles, t=np.zeros((num_rotations, 2), dtype=F))
    c = get_grid_center_transform(grid)
    return c @ rotated @ c.inv


def sample_query_templates(features: np.ndarray, valid: np.ndarray, num_rotations: int, grid: grids.Grid2D):
    """:37-69.  features [H,W,D], valid [H,W] -> templates [R,H,W,D], t_valid [R,H,W]."""
    assert grid.extent[0] == grid.extent[1] and num_rotations % 4 == 0
    tfm = template_transforms(num_rotations, grid)
    grid_xy = grid.index_to_xyz(grid.grid_index()).reshape(-1, 2)
    nq = num_rotations // 4
    quarter, t_valid = [], []
    for r in range(nq):
        one = geometry.Transform2D(angle=tfm.angle[r], t=tfm.t[r])
        xy = one.transform(grid_xy)
        uv = (xy / F(grid.cell_size)).astype(F)
        q, v = grids.interpolate_nd(features, uv, valid)
        quarter.append(np.where(v[..., None], q, F(0)))
        t_valid.append(v)
    quarter = np.stack(quarter).reshape(nq, *grid.extent, features.shape[-1])
    t_valid = np.stack(t_valid).reshape(nq, *grid.extent)
    templates = np.concatenate([np.rot90(quarter, k, axes=(2, 1)) for k in range(4)], 0)
    t_valid = np.concatenate([np.rot90(t_valid, k, axes=(2, 1)) for k in range(4)], 0)
    return templates.astype(F), t_valid


def template_matching(q: np.ndarray, q_valid: np.ndarray, m: np.ndarray, m_valid: np.ndarray,
                      min_overlap: Optional[float] = 0.05) -> np.ndarray:
    """:72-104 with do_padding=True.  q [R,H,W,D], m [H,W,D] -> scores [R,2H-1,2W-1] fp32.

    `convolve(q[:, ::-1, ::-1], m_pad, 'valid')` summed over channels is the cross-correlation
    S_r[u,v] = sum_ijd q_r[i,j,d] m_pad[u+i,v+j,d] (torch conv2d is a correlation).  The overlap
    count convolves the UN-flipped q_valid with the zero-padded m_valid.
    """
    R, H, W, D = q.shape
    mt = torch.from_numpy(np.ascontiguousarray(m, dtype=F)).permute(2, 0, 1)[None]
    m_pad = Fnn.pad(mt, (W - 1, W - 1, H - 1, H - 1), mode="replicate")  # jnp.pad(mode='edge') :83-85
    qt = torch.from_numpy(np.ascontiguousarray(q, dtype=F)).permute(0, 3, 1, 2)
    scores = Fnn.conv2d(m_pad, qt)[0].numpy()  # [R, 2H-1, 2W-1]
    if min_overlap is not None:
        mv = torch.from_numpy(m_valid.astype(F))[None, None]
        mv_pad = Fnn.pad(mv, (W - 1, W - 1, H - 1, H - 1))  # constant 0 :94-96
        qv = torch.from_numpy(np.ascontiguousarray(q_valid[:, ::-1, ::-1]).astype(F))[:, None]  # true convolution
        num_valid = Fnn.conv2d(mv_pad, qv)[0].numpy()
        thr = F(min_overlap * math.prod(q_valid.shape[-2:]))
        scores = np.where(num_valid > thr, scores, -np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        scores = scores / q_valid.sum((-1, -2), keepdims=True).astype(F)  # :103
    return scores.astype(F)


def exhaustive_pose_voting(feats_q, valid_q, feats_m, valid_m, num_rotations: int, grid: grids.Grid2D,
                           conf_q: Optional[np.ndarray] = None) -> np.ndarray:
    """:107-124."""
    if conf_q is not None:
        feats_q = feats_q * conf_q[..., None]
    templates, t_valid = sample_query_templates(feats_q, valid_q, num_rotations, grid)
    return template_matching(templates, t_valid, feats_m, valid_m)


def exhaustive_index_to_tfm(index: np.ndarray, grid: grids.Grid2D, num_rotations: int) -> geometry.Transform2D:
    """:127-136."""
    xy_cell = ((index[1:] - np.array(grid.extent) + 1 + 0.5) * grid.cell_size).astype(F)
    angle = F(index[0] * 2 * np.pi / num_rotations)
    m_t_q_center = geometry.Transform2D(angle=np.asarray(-angle, dtype=F), t=xy_cell)
    c = get_grid_center_transform(grid)
    return c @ m_t_q_center @ c.inv


def exhaustive_tfm_to_index(m_t_q_corner: geometry.Transform2D, grid: grids.Grid2D, num_rotations: int):
    """:139-149."""
    c = get_grid_center_transform(grid)
    m = c.inv @ m_t_q_corner @ c
    k = (-m.angle / (np.pi * 2) % 1) * num_rotations
    ij = (m.t / grid.cell_size) + np.array(grid.extent) - 1.5
    return np.concatenate([np.asarray(k)[..., None], ij], -1)
