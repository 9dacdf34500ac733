"""B200 implementation of `snap/models/pose_exhaustive_voting.py` (exhaustive (x, y, theta) voting)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops, types

F = np.float32


def _compose(a_angle, a_t, b_angle, b_t):
    """Transform2D.compose (`snap/utils/geometry.py:142-145`): angle sum, t = a.t + R(a) b.t (fp32, j order)."""
    cos, sin = np.cos(a_angle).astype(F), np.sin(a_angle).astype(F)
    R = np.stack([cos, -sin, sin, cos], -1).reshape(*np.shape(a_angle), 2, 2).astype(F)
    s = (R[..., 0] * b_t[..., None, 0]).astype(F)
    s = (s + (R[..., 1] * b_t[..., None, 1]).astype(F)).astype(F)
    return (a_angle + b_angle).astype(F), (a_t + s).astype(F)


def _inv(angle, t):
    """Transform2D.inv (`geometry.py:126-130`)."""
    cos, sin = np.cos(angle).astype(F), np.sin(angle).astype(F)
    R_inv = np.stack([cos, sin, -sin, cos], -1).reshape(*np.shape(angle), 2, 2).astype(F)
    s = (R_inv[..., 0] * t[..., None, 0]).astype(F)
    s = (s + (R_inv[..., 1] * t[..., None, 1]).astype(F)).astype(F)
    return (-angle).astype(F), (-s).astype(F)


def template_rotation_params(num_rotations: int, grid: types.Grid2D) -> np.ndarray:
    """(cos, sin, tx, ty) of templates_t_grid = corner_t_center @ rotated_t_grid @ corner_t_center.inv for
    the first R/4 rotations (`pose_exhaustive_voting.py:45-54`), fp32 host math."""
    angles = np.linspace(0, np.pi * 2, num_rotations, endpoint=False).astype(F)
    c_angle = np.asarray(0, dtype=F)
    c_t = (np.asarray(grid.extent_meters) / 2).astype(F)
    ci_angle, ci_t = _inv(c_angle, c_t)
    a1, t1 = _compose(c_angle, c_t, angles, np.zeros((num_rotations, 2), dtype=F))
    a2, t2 = _compose(a1, t1, ci_angle, ci_t)
    nq = num_rotations // 4
    return np.stack([np.cos(a2[:nq]).astype(F), np.sin(a2[:nq]).astype(F), t2[:nq, 0], t2[:nq, 1]], -1).astype(F)


_CENTERS: dict = {}


def sample_query_templates(features: torch.Tensor, valid: torch.Tensor, num_rotations: int, grid: types.Grid2D,
                           conf_q: Optional[torch.Tensor] = None):
    """`:37-69`, batched: features bf16 [B,G,G,D], valid u8 [B,G,G] -> templates, t_valid u8 [B,R,G,G].

    Templates are returned CELL-MAJOR, bf16 [B, G, G, RP, D] with RP = R rounded up to a multiple of 48 (zero rows):
    `templates[b, i, j, r]` is the reference's `templates[r, i, j]`; this is the B operand layout of the
    correlation GEMM.  Use `templates_to_reference_layout` to get [B,R,G,G,D]."""
    if grid.extent[0] != grid.extent[1] or num_rotations % 4:
        raise ValueError("exhaustive voting needs a square grid and num_rotations % 4 == 0")
    B, G, _, D = features.shape
    dev = features.device
    rot = template_rotation_params(num_rotations, grid)
    key = (str(dev), grid.extent[0], float(grid.cell_size))
    if key not in _CENTERS:      # uploaded once per device and grid (no H2D copy inside a captured graph)
        _CENTERS[key] = torch.from_numpy(grid.cell_centers(0)).to(dev)
    centers = _CENTERS[key]
    templates = torch.empty((B, G, G, ops.xcorr_padded_rotations(num_rotations), D), dtype=torch.bfloat16, device=dev)
    t_valid = torch.empty((B, num_rotations, G, G), dtype=torch.uint8, device=dev)
    ops.rot_templates(features.contiguous(), valid.contiguous(), conf_q, rot, centers, float(F(grid.cell_size)),
                      num_rotations, templates, t_valid)
    return templates, t_valid


def templates_to_reference_layout(templates: torch.Tensor, num_rotations: int) -> torch.Tensor:
    """cell-major [B,G,G,RP,D] -> the reference's [B,R,G,G,D]."""
    return templates[:, :, :, :num_rotations].permute(0, 3, 1, 2, 4).contiguous()


def template_matching(q: torch.Tensor, q_valid: torch.Tensor, m: torch.Tensor, m_valid: torch.Tensor,
                      min_overlap: Optional[float] = 0.05, kernel: str = "auto") -> torch.Tensor:
    """`:72-104` (do_padding=True), batched: q = cell-major templates bf16 [B,G,G,RP,D], q_valid u8 [B,R,G,G],
    m bf16 [B,G,G,D] -> f32 [B,R,2G-1,2G-1]."""
    B, G, _, RP, D = q.shape
    R = q_valid.shape[1]
    dev = q.device
    U = 2 * G - 1
    m_pad = torch.empty((B, 3 * G - 2, ops.xcorr_padded_cols(G), D), dtype=torch.bfloat16, device=dev)
    ops.xcorr_pad_map(m.contiguous(), m_pad)
    cnt = torch.empty((B, R, U, U), dtype=torch.float32, device=dev)
    den = torch.empty((B, R), dtype=torch.float32, device=dev)
    ops.xcorr_count(q_valid.contiguous(), m_valid.contiguous(), cnt, den)
    scores = torch.empty((B, R, U, U), dtype=torch.float32, device=dev)
    thr = float(F(min_overlap * G * G)) if min_overlap is not None else 0.0
    # "rows": map-row-major stacked-template MMAs (fastest); "sw": sliding window, N = 48; "gemm": segmented GEMM
    if kernel == "auto":
        kernel = "rows" if ops.xcorr_rows_supported(R, G) else ("sw" if ops.xcorr_sw_supported(R, G) else "gemm")
    fn = {"rows": ops.xcorr_scores_rows, "sw": ops.xcorr_scores_sw, "gemm": ops.xcorr_scores}[kernel]
    fn(q.contiguous(), m_pad, cnt if min_overlap is not None else None, den, thr, scores)
    return scores


def exhaustive_pose_voting(plane_q: types.FeaturePlane, plane_map: types.FeaturePlane, num_rotations: int,
                           grid: types.Grid2D, conf_q: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`:107-124`; planes carry a leading batch axis (the reference vmaps this function)."""
    templates, t_valid = sample_query_templates(plane_q.features, plane_q.valid, num_rotations, grid, conf_q)
    return template_matching(templates, t_valid, plane_map.features, plane_map.valid)


def exhaustive_index_to_tfm(index: np.ndarray, grid: types.Grid2D, num_rotations: int):
    """`:127-136`: pose-volume index -> (angle, t) of m_t_q w.r.t. the grid corner."""
    xy_cell = ((index[1:] - np.array(grid.extent) + 1 + 0.5) * grid.cell_size).astype(F)
    angle = F(index[0] * 2 * np.pi / num_rotations)
    c_angle, c_t = np.asarray(0, dtype=F), (np.asarray(grid.extent_meters) / 2).astype(F)
    a1, t1 = _compose(c_angle, c_t, np.asarray(-angle, dtype=F), xy_cell)
    return _compose(a1, t1, *_inv(c_angle, c_t))


def exhaustive_tfm_to_index(angle, t, grid: types.Grid2D, num_rotations: int) -> np.ndarray:
    """`:139-149`."""
    c_angle, c_t = np.asarray(0, dtype=F), (np.asarray(grid.extent_meters) / 2).astype(F)
    ci = _inv(c_angle, c_t)
    a1, t1 = _compose(*ci, np.asarray(angle, dtype=F), np.asarray(t, dtype=F))
    a2, t2 = _compose(a1, t1, c_angle, c_t)
    k = (-a2 / (np.pi * 2) % 1) * num_rotations
    ij = (t2 / grid.cell_size) + np.array(grid.extent) - 1.5
    return np.concatenate([np.asarray(k)[..., None], ij], -1)
