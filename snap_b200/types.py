"""Containers mirroring `snap/models/types.py`, `snap/utils/grids.py` and `snap/utils/geometry.py`
(the data side of the module boundary).  Host-side fields are NumPy fp32; device features are torch."""
from __future__ import annotations

import dataclasses
from typing import Any, List, Optional, Tuple

import numpy as np

F = np.float32


@dataclasses.dataclass
class FeatureVolume:  # snap/models/types.py:24-29
    features: Any
    valid: Optional[Any] = None


@dataclasses.dataclass
class FeaturePlane:  # snap/models/types.py:32-37
    features: Any
    valid: Optional[Any] = None


@dataclasses.dataclass
class FeatureImagePyramid:  # snap/models/types.py:40-45
    features: List[Any]
    strides: List[Any]
    # B200 extra: the un-cropped, contiguous [n, Hs, Ws, C] buffers the cropped `features` view into
    uncropped: Optional[List[Any]] = None


@dataclasses.dataclass(frozen=True)
class Grid2D:  # snap/utils/grids.py:33-93
    extent: Tuple[int, int]
    cell_size: float

    def cell_centers(self, axis: int) -> np.ndarray:
        """index_to_xyz (grids.py:62-63) along one axis: (idx + 0.5) * cell_size in fp32."""
        idx = np.arange(self.extent[axis], dtype=np.int32)
        return ((idx.astype(F) + F(0.5)) * F(self.cell_size)).astype(F)

    @property
    def extent_meters(self) -> np.ndarray:
        return np.asarray(self.extent) * self.cell_size


@dataclasses.dataclass
class Transform3D:  # snap/utils/geometry.py:36-84
    R: np.ndarray  # [..., 3, 3]
    t: np.ndarray  # [..., 3]

    @property
    def inv(self) -> "Transform3D":  # geometry.py:52-56, fixed fp32 summation order j = 0, 1, 2
        R_inv = np.ascontiguousarray(np.swapaxes(self.R, -1, -2)).astype(F)
        s = (R_inv[..., 0] * self.t[..., None, 0]).astype(F)
        s = (s + (R_inv[..., 1] * self.t[..., None, 1]).astype(F)).astype(F)
        s = (s + (R_inv[..., 2] * self.t[..., None, 2]).astype(F)).astype(F)
        return Transform3D(R=R_inv, t=(-s).astype(F))


@dataclasses.dataclass
class Camera:  # snap/utils/geometry.py:160-221 (pinhole); fields [..., 2] in (x, y) order
    wh: np.ndarray
    f: np.ndarray
    c: np.ndarray

    def scale(self, scale) -> "Camera":  # geometry.py:179-183
        s = np.asarray(scale, dtype=F)
        return dataclasses.replace(self, wh=(self.wh * s).astype(F), f=(self.f * s).astype(F),
                                   c=(self.c * s).astype(F))


@dataclasses.dataclass
class FisheyeCamera(Camera):  # snap/utils/geometry.py:224-280
    k_radial: np.ndarray = None  # [..., 3]
    max_fov: np.ndarray = None   # [...], radians
