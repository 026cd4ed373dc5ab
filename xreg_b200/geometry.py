"""Host-side geometry types mirroring the reference's (float32 throughout,
lib/common/xregCommon.h:47,65,128): CameraModel, Volume, SE(3) helpers.

These only prepare the POD inputs of the C ABI; no image-sized work happens here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _lib

# CameraModel::CameraCoordFrame (lib/transforms/xregPerspectiveXform.h:128-133)
kORIGIN_AT_FOCAL_PT_DET_POS_Z = 0
kORIGIN_AT_FOCAL_PT_DET_NEG_Z = 1
kORIGIN_ON_DETECTOR = 2

f32 = np.float32


def _inv3(m: np.ndarray) -> np.ndarray:
    """3x3 f32 inverse by cofactors (same shape as Eigen's fixed-size inverse)."""
    m = np.asarray(m, dtype=f32)
    c = np.empty((3, 3), dtype=f32)
    for i in range(3):
        for j in range(3):
            i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
            c[i, j] = f32(m[i1, j1] * m[i2, j2]) - f32(m[i1, j2] * m[i2, j1])
    det = f32(f32(f32(c[0, 0] * m[0, 0]) + f32(c[1, 0] * m[1, 0])) + f32(c[2, 0] * m[2, 0]))
    return (c.T * (f32(1) / det)).astype(f32)


def se3_inv(T: np.ndarray) -> np.ndarray:
    """SE3Inv (lib/transforms/xregRigidUtils.cpp:29-38)"""
    T = np.asarray(T, dtype=f32)
    out = np.eye(4, dtype=f32)
    out[:3, :3] = T[:3, :3].T
    # -1 * R^T * t with the products summed left to right in f32 (a BLAS matmul may fuse or reorder: 1 ulp off)
    for r in range(3):
        out[r, 3] = -_dot3(out[r, 0], out[r, 1], out[r, 2], T[0, 3], T[1, 3], T[2, 3])
    return out


def _dot3(a0, a1, a2, b0, b1, b2) -> np.float32:
    """(a0 b0 + a1 b1) + a2 b2 in f32, each operation rounded (Eigen's small fixed-size products, no FMA)."""
    return f32(f32(f32(f32(a0) * f32(b0)) + f32(f32(a1) * f32(b1))) + f32(f32(a2) * f32(b2)))


def skew(w: Sequence[float]) -> np.ndarray:
    x, y, z = (f32(v) for v in w)
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=f32)


def exp_se3(x: Sequence[float]) -> np.ndarray:
    """ExpSE3(Pt6) (lib/transforms/xregRigidUtils.cpp:40-85): x = [w_x, w_y, w_z, v_x, v_y, v_z]."""
    x = np.asarray(x, dtype=f32)
    W = skew(x[:3])
    v = x[3:6]
    T = np.eye(4, dtype=f32)
    theta = f32(np.linalg.norm(x[:3]))
    if theta > 1.0e-14:
        Wu = W / theta
        R = np.eye(3, dtype=f32) + f32(np.sin(theta)) * Wu + f32(1 - np.cos(theta)) * (Wu @ Wu)
        th2 = theta * theta
        A = np.eye(3, dtype=f32) + f32((1 - np.cos(theta)) / th2) * W + f32((theta - np.sin(theta)) / (theta * th2)) * (W @ W)
        T[:3, :3] = R
        T[:3, 3] = A @ v
    else:
        T[:3, 3] = v
    return T.astype(f32)


def to12(T: np.ndarray) -> np.ndarray:
    """Top 3x4 of a 4x4 (or an Nx4x4 stack) as row-major 12-vectors."""
    T = np.asarray(T, dtype=f32)
    if T.ndim == 2:
        return np.ascontiguousarray(T[:3, :].reshape(12))
    return np.ascontiguousarray(T[:, :3, :].reshape(-1, 12))


@dataclass
class CameraModel:
    """CameraModel (lib/transforms/xregPerspectiveXform.h:108-273)."""

    coord_frame_type: int = kORIGIN_AT_FOCAL_PT_DET_NEG_Z
    intrins: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=f32))
    intrins_inv: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=f32))
    extrins: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=f32))
    extrins_inv: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=f32))
    pinhole_pt: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=f32))
    focal_len: float = 0.0
    num_det_rows: int = 0
    num_det_cols: int = 0
    det_row_spacing: float = 0.0
    det_col_spacing: float = 0.0

    def setup(self, focal_len: float, nr: int, nc: int, rs: float, cs: float) -> "CameraModel":
        """setup(focal_len, nr, nc, rs, cs) with MakeNaiveIntrins (xregPerspectiveXform.cpp:200-254)."""
        if not (focal_len > 1.0e-8 and nr and nc and rs > 1.0e-8 and cs > 1.0e-8):
            raise _lib.XregError("CameraModel.setup: invalid arguments")
        self.focal_len = float(f32(focal_len))
        self.num_det_rows, self.num_det_cols = int(nr), int(nc)
        self.det_row_spacing, self.det_col_spacing = float(f32(rs)), float(f32(cs))
        K = np.eye(3, dtype=f32)
        K[0, 0] = f32(focal_len) / f32(cs)
        K[1, 1] = f32(focal_len) / f32(rs)
        if self.coord_frame_type == kORIGIN_AT_FOCAL_PT_DET_NEG_Z:
            K[0, 0] *= -1
            K[1, 1] *= -1
        K[0, 2] = f32((nc - 1) * 0.5)
        K[1, 2] = f32((nr - 1) * 0.5)
        self.intrins = K
        self.intrins_inv = _inv3(K)
        self.extrins = np.eye(4, dtype=f32)
        self.extrins_inv = np.eye(4, dtype=f32)
        self.pinhole_pt = np.zeros(3, dtype=f32)
        return self

    def setup_intrins_extrins(self, intrins, extrins, nr: int, nc: int, rs: float, cs: float) -> "CameraModel":
        """setup(intrins, extrins, ...) (xregPerspectiveXform.cpp:302-334)."""
        self.num_det_rows, self.num_det_cols = int(nr), int(nc)
        self.det_row_spacing, self.det_col_spacing = float(f32(rs)), float(f32(cs))
        self.intrins = np.asarray(intrins, dtype=f32).copy()
        self.intrins_inv = _inv3(self.intrins)
        yps = self.det_col_spacing if rs < 0 else self.det_row_spacing
        self.focal_len = float((abs(f32(self.intrins[0, 0] * f32(cs))) + abs(f32(self.intrins[1, 1] * f32(yps)))) / f32(2))
        self.extrins = np.asarray(extrins, dtype=f32).copy()
        self.extrins_inv = se3_inv(self.extrins)
        if self.coord_frame_type in (kORIGIN_AT_FOCAL_PT_DET_POS_Z, kORIGIN_AT_FOCAL_PT_DET_NEG_Z):
            self.pinhole_pt = self.extrins_inv[:3, 3].copy()
        else:
            ei = self.extrins_inv
            self.pinhole_pt = np.array([f32(_dot3(ei[r, 0], ei[r, 1], ei[r, 2], 0.0, 0.0, self.focal_len) + ei[r, 3])
                                        for r in range(3)], dtype=f32)
        return self

    def to_xrc(self) -> _lib.XrcCam:
        s = _lib.XrcCam()
        s.rows, s.cols = self.num_det_rows, self.num_det_cols
        s.intrins_inv[:] = [float(v) for v in np.asarray(self.intrins_inv, dtype=f32).reshape(9)]
        s.extrins_inv[:] = [float(v) for v in np.asarray(self.extrins_inv, dtype=f32)[:3, :].reshape(12)]
        s.pinhole[:] = [float(v) for v in np.asarray(self.pinhole_pt, dtype=f32).reshape(3)]
        s.focal_len = float(self.focal_len)
        s.frame_type = int(self.coord_frame_type)
        return s


def downsample_camera_model(src: CameraModel, ds_factor: float, force_even_dims: bool = False) -> CameraModel:
    """DownsampleCameraModel (lib/transforms/xregPerspectiveXform.cpp:654-688)."""
    dst = CameraModel(coord_frame_type=src.coord_frame_type)
    K = src.intrins.astype(f32).copy()
    K[0, 0] *= f32(ds_factor)
    K[1, 1] *= f32(ds_factor)
    K[0, 2] *= f32(ds_factor)
    K[1, 2] *= f32(ds_factor)
    nr = int(np.floor(src.num_det_rows * f32(ds_factor) + 0.5))  # std::lround
    nc = int(np.floor(src.num_det_cols * f32(ds_factor) + 0.5))
    if force_even_dims:
        nr -= nr % 2
        nc -= nc % 2
    return dst.setup_intrins_extrins(K, src.extrins, nr, nc, f32(src.det_row_spacing) / f32(ds_factor),
                                     f32(src.det_col_spacing) / f32(ds_factor))


@dataclass
class Volume:
    """Stand-in for itk::Image<float,3>: x-fastest voxels plus ITK metadata (doubles)."""

    data: np.ndarray  # (nz, ny, nx) float32, C-contiguous
    spacing: Sequence[float] = (1.0, 1.0, 1.0)
    origin: Sequence[float] = (0.0, 0.0, 0.0)
    direction: np.ndarray = field(default_factory=lambda: np.eye(3))

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data, dtype=f32)
        if self.data.ndim != 3:
            raise _lib.XregError("Volume: expected a 3-D array (nz, ny, nx)")

    @property
    def dims(self):
        nz, ny, nx = self.data.shape
        return (nx, ny, nz)

    def idx_to_phys(self) -> np.ndarray:
        """ITKImagePhysicalPointTransformsAsEigen (lib/itk/xregITKBasicImageUtils.h:131-168):
        M[r][c] = float(Dir[r][c] * spacing[c]), t[r] = float(origin[r])."""
        D = np.asarray(self.direction, dtype=np.float64)
        sp = np.asarray(self.spacing, dtype=np.float64)
        out = np.zeros((3, 4), dtype=f32)
        out[:, :3] = (D * sp[None, :]).astype(f32)
        out[:, 3] = np.asarray(self.origin, dtype=np.float64).astype(f32)
        return out.reshape(12)
