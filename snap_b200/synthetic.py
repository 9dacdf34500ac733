"""Seeded synthetic StreetView(+aerial) tiles of SURVEY.md §8(d): uniform images, pinhole 72-degree
cameras on a street line through the grid centre, side-looking, gravity aligned."""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import types

F = np.float32


def _rot_cam(yaw: float) -> np.ndarray:
    """Columns = camera axes (x right, y down, z forward) in the scene frame (z up); yaw about z."""
    fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
    down = np.array([0.0, 0.0, -1.0])
    right = np.cross(down, fwd)
    return np.stack([right, down, fwd], axis=1)


def make_tile(tile_id: int, num_views: int, image_hw, grid_side: int, cell_size: float = 0.2,
              aerial: bool = False, batch: int = 1, fisheye: bool = False, spacing: float = 3.0,
              same_side: bool = False) -> Dict:
    """Returns the reference's batch dict (`snap/data/loader.py:89-110`) with NumPy leaves:
    'images' f32 [B,V,H,W,3], 'camera' Camera [B,V], 'T_view2scene' Transform3D [B,V], optional
    'rasters': {'rgb' [B,G,G,3]}.  `spacing` (m between consecutive cameras) and `same_side` (all cameras look to
    the same side of the street instead of alternating) control how many views overlap on a voxel: the defaults
    give the sparse street layout of SURVEY §8(d), small spacing + same_side the dense many-view case that the
    top-k view selection exists for."""
    H, W = image_hw
    ext = grid_side * cell_size
    images, Rs, ts, rgbs = [], [], [], []
    for b in range(batch):
        rng = np.random.default_rng(1234 + tile_id * 1000 + b)
        images.append(rng.random((num_views, H, W, 3), dtype=F))
        R_b, t_b = [], []
        for v in range(num_views):
            along = (v - (num_views - 1) / 2) * spacing + rng.uniform(-0.5, 0.5) * spacing / 3.0
            pos = np.array([ext / 2 + along, ext / 2 + rng.uniform(-0.3, 0.3), 2.5 + rng.uniform(-0.2, 0.2)])
            yaw = (np.pi / 2 if (v % 2 == 0 or same_side) else -np.pi / 2) + np.deg2rad(rng.uniform(-10, 10))
            R_b.append(_rot_cam(yaw))
            t_b.append(pos)
        Rs.append(np.stack(R_b)); ts.append(np.stack(t_b))
        if aerial:
            rgbs.append(rng.random((grid_side, grid_side, 3), dtype=F))
    f = (W / 2) / np.tan(np.deg2rad(36.0))
    shape = (batch, num_views, 2)
    cam_kw = dict(wh=np.broadcast_to(np.array([W, H], F), shape).copy(),
                  f=np.broadcast_to(np.array([f, f], F), shape).copy(),
                  c=np.broadcast_to(np.array([W / 2, H / 2], F), shape).copy())
    if fisheye:
        camera = types.FisheyeCamera(**cam_kw,
                                     k_radial=np.broadcast_to(np.array([-0.03, 0.005, 0.0], F), (batch, num_views, 3)).copy(),
                                     max_fov=np.full((batch, num_views), np.deg2rad(115.0), F))
    else:
        camera = types.Camera(**cam_kw)
    data = {"images": np.stack(images), "camera": camera,
            "T_view2scene": types.Transform3D(R=np.stack(Rs).astype(F), t=np.stack(ts).astype(F))}
    if aerial:
        data["rasters"] = {"rgb": np.stack(rgbs)}
    return data
