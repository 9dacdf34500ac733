"""Restatement of snap/utils/grids.py (reference file:line cited per function). Test infrastructure."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np


@dataclasses.dataclass(frozen=True)
class Grid2D:
    """snap/utils/grids.py:33-93 (GridND/Grid2D)."""
    extent: Tuple[int, int]
    cell_size: float

    def index_to_xyz(self, idx: np.ndarray) -> np.ndarray:  # grids.py:62-63
        return ((idx.astype(np.float32) + np.float32(0.5)) * np.float32(self.cell_size)).astype(np.float32)

    def grid_index(self) -> np.ndarray:  # grids.py:87-89
        g = np.mgrid[tuple(slice(None, e) for e in self.extent)]
        return np.moveaxis(g, 0, -1).astype(np.int32)

    @property
    def extent_meters(self) -> np.ndarray:  # grids.py:77-79
        return np.asarray(self.extent) * self.cell_size


def map_coordinates_linear_nearest(array: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """jax.scipy.ndimage.map_coordinates(order=1, mode='nearest') per SURVEY Appendix A.1.

    array [d0..dn-1], coords [n, K] -> [K].  Per dim lo=floor(c), w_hi=c-lo, taps lo / lo+1 are
    index-clamped; corners accumulated in itertools.product order (lower first).
    """
    n = array.ndim
    K = coords.shape[1]
    out = np.zeros(K, dtype=np.float32)
    lo = np.floor(coords).astype(np.int64)
    w_hi = (coords - lo).astype(np.float32)
    w_lo = (np.float32(1.0) - w_hi).astype(np.float32)
    for corner in np.ndindex(*([2] * n)):
        w = np.ones(K, dtype=np.float32)
        idx = []
        for d in range(n):
            tap = lo[d] + corner[d]
            idx.append(np.clip(tap, 0, array.shape[d] - 1))
            w = w * (w_hi[d] if corner[d] else w_lo[d])
        out = out + w * array[tuple(idx)].astype(np.float32)
    return out


def interpolate_nd(array: np.ndarray, points: np.ndarray, valid_array: Optional[np.ndarray] = None):
    """snap/utils/grids.py:116-137.  array [..., D], points [K, N] -> values [K, D], valid [K]."""
    size = np.asarray(array.shape[:-1])
    valid = np.all((points >= 0) & (points < size), -1)  # grids.py:126 (un-shifted point)
    pts = np.moveaxis(points.astype(np.float32) - np.float32(0.5), -1, 0)  # grids.py:129
    values = np.stack([map_coordinates_linear_nearest(array[..., d], pts) for d in range(array.shape[-1])], -1)
    if valid_array is not None:  # grids.py:131-136: 0/NaN mask, NaN if ANY tap invalid (even weight 0)
        with np.errstate(invalid="ignore"):
            nan_mask = np.where(valid_array, np.float32(0), np.float32(np.nan)).astype(np.float32)
            nan_pts = map_coordinates_linear_nearest(nan_mask, pts)
        valid = valid & ~np.isnan(nan_pts)
    return values.astype(array.dtype), valid
