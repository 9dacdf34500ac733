"""Restatement of snap/utils/geometry.py (struct-of-arrays, NumPy fp32). Test infrastructure.

Arithmetic is written as explicit elementwise fp32 multiply/add in a FIXED order (no FMA, no BLAS)
so that the CUDA kernels can reproduce visibility masks and tap indices bit for bit.
"""
from __future__ import annotations

import dataclasses

import numpy as np

F = np.float32


@dataclasses.dataclass
class Transform3D:
    """snap/utils/geometry.py:36-84.  R [...,3,3], t [...,3]."""
    R: np.ndarray
    t: np.ndarray

    @property
    def inv(self) -> "Transform3D":  # geometry.py:52-56
        R_inv = np.swapaxes(self.R, -1, -2)
        t_inv = -_matvec3(R_inv, self.t)
        return Transform3D(R=R_inv.astype(F), t=t_inv.astype(F))

    def transform(self, p3d: np.ndarray) -> np.ndarray:  # geometry.py:67-69
        """p3d [..., n, 3] -> t + R p  (sum over j in order j=0,1,2)."""
        R = self.R[..., None, :, :]
        p = p3d[..., None, :]
        s = (R[..., 0] * p[..., 0]).astype(F)
        s = (s + (R[..., 1] * p[..., 1]).astype(F)).astype(F)
        s = (s + (R[..., 2] * p[..., 2]).astype(F)).astype(F)
        return (self.t[..., None, :] + s).astype(F)


def _matvec3(R: np.ndarray, v: np.ndarray) -> np.ndarray:
    s = (R[..., 0] * v[..., None, 0]).astype(F)
    s = (s + (R[..., 1] * v[..., None, 1]).astype(F)).astype(F)
    s = (s + (R[..., 2] * v[..., None, 2]).astype(F)).astype(F)
    return s


@dataclasses.dataclass
class Camera:
    """Pinhole camera, snap/utils/geometry.py:160-221.  wh, f, c: [..., 2] (x, y order)."""
    wh: np.ndarray
    f: np.ndarray
    c: np.ndarray
    eps = 1e-3

    def scale(self, scale: np.ndarray) -> "Camera":  # geometry.py:179-183
        s = np.asarray(scale, dtype=F)
        return Camera(wh=(self.wh * s).astype(F), f=(self.f * s).astype(F), c=(self.c * s).astype(F))

    def in_image(self, p2d: np.ndarray) -> np.ndarray:  # geometry.py:193-196
        return np.all((p2d >= 0) & (p2d < self.wh[..., None, :]), -1)

    def project(self, p3d: np.ndarray):  # geometry.py:198-205
        z = p3d[..., -1]
        valid = z >= F(self.eps)
        z = np.maximum(z, F(self.eps))[..., None]
        p2d = (p3d[..., :-1] / z).astype(F)
        return p2d, valid

    def denormalize(self, p2d: np.ndarray) -> np.ndarray:  # geometry.py:207-210
        return ((p2d * self.f[..., None, :]).astype(F) + self.c[..., None, :]).astype(F)

    def world2image(self, p3d: np.ndarray):  # geometry.py:216-221
        p2d, visible = self.project(p3d)
        p2d = self.denormalize(p2d)
        valid = visible & self.in_image(p2d)
        return p2d, valid


@dataclasses.dataclass
class FisheyeCamera(Camera):
    """snap/utils/geometry.py:224-280."""
    k_radial: np.ndarray = None
    max_fov: np.ndarray = None

    def scale(self, scale):  # geometry.py:250-258
        s = np.asarray(scale, dtype=F)
        return FisheyeCamera(wh=(self.wh * s).astype(F), f=(self.f * s).astype(F), c=(self.c * s).astype(F),
                             k_radial=self.k_radial, max_fov=self.max_fov)

    def distort_points(self, p2d):  # geometry.py:260-272
        eps2 = F(self.eps) ** 2
        radius2 = np.sum(p2d ** 2, axis=-1).astype(F)
        in_center = radius2 < eps2
        radius = np.sqrt(np.where(in_center, eps2, radius2)).astype(F)
        theta = np.arctan(radius).astype(F)
        theta2 = (theta ** 2).astype(F)
        offset = sum(self.k_radial[..., None, i] * theta2 ** (i + 1) for i in range(3))
        dist = ((offset + 1) * theta / radius).astype(F)
        dist = np.where(in_center, F(1.0), dist)
        p2d_dist = (p2d * dist[..., None]).astype(F)
        valid = in_center | ((radius < np.tan(0.5 * self.max_fov)[..., None]) & (dist > 0))
        return p2d_dist, valid

    def world2image(self, p3d):  # geometry.py:274-280
        p2d, visible = self.project(p3d)
        p2d, valid = self.distort_points(p2d)
        p2d = self.denormalize(p2d)
        valid = visible & valid & self.in_image(p2d)
        return p2d, valid


@dataclasses.dataclass
class Transform2D:
    """snap/utils/geometry.py:87-154 (angle, t)."""
    angle: np.ndarray
    t: np.ndarray

    @property
    def R(self):  # geometry.py:113-118
        cos, sin = np.cos(self.angle).astype(F), np.sin(self.angle).astype(F)
        return np.stack([cos, -sin, sin, cos], -1).reshape(*np.shape(self.angle), 2, 2).astype(F)

    @property
    def inv(self):  # geometry.py:126-130
        R_inv = np.swapaxes(self.R, -1, -2)
        t_inv = -_matvec2(R_inv, self.t)
        return Transform2D(angle=(-self.angle).astype(F), t=t_inv.astype(F))

    def transform(self, points):  # geometry.py:138-140
        R = self.R[..., None, :, :]
        p = points[..., None, :]
        s = (R[..., 0] * p[..., 0]).astype(F)
        s = (s + (R[..., 1] * p[..., 1]).astype(F)).astype(F)
        return (self.t[..., None, :] + s).astype(F)

    def compose(self, other):  # geometry.py:142-145
        angle = (self.angle + other.angle).astype(F)
        t = (self.t + _matvec2(self.R, other.t)).astype(F)
        return Transform2D(angle=angle, t=t)

    def __matmul__(self, other):
        if isinstance(other, Transform2D):
            return self.compose(other)
        return self.transform(other)


def _matvec2(R, v):
    s = (R[..., 0] * v[..., None, 0]).astype(F)
    s = (s + (R[..., 1] * v[..., None, 1]).astype(F)).astype(F)
    return s
