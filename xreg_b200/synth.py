"""Deterministic synthetic inputs of the benchmark configurations:
CT-like volume, C-arm camera, nominal pose, CMA-ES-like pose population.
Used by tests/ and bench.py; numpy only.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from .geometry import (CameraModel, Volume, downsample_camera_model, exp_se3, f32,
                       kORIGIN_AT_FOCAL_PT_DET_NEG_Z)

SEED = 20211009


def make_volume(nx: int, ny: int, nz: int, spacing: Sequence[float] = (1.0, 1.0, 1.0), seed: int = SEED,
                n_bones: int = 12, margin: int = 8) -> Volume:
    """Linear-attenuation phantom: air, ellipsoidal soft-tissue body (0.02 mm^-1 plus
    smooth low-frequency texture), random ellipsoidal bones (0.05 mm^-1)."""
    rng = np.random.default_rng(seed)
    coarse = rng.uniform(0.0, 0.01, size=(16, 16, 16)).astype(np.float64)
    bones = []
    for _ in range(n_bones):
        c = rng.uniform(-0.3, 0.3, size=3)
        r = rng.uniform(0.04, 0.12, size=3)
        bones.append((c, r))

    # normalised coordinates in [-1, 1] across the volume extent
    xs = (np.arange(nx) - (nx - 1) / 2.0) / (nx / 2.0)
    ys = (np.arange(ny) - (ny - 1) / 2.0) / (ny / 2.0)
    zs = (np.arange(nz) - (nz - 1) / 2.0) / (nz / 2.0)
    ax = 0.9 * (1.0 - 2.0 * margin / max(nx, 2 * margin + 1))
    ay = 0.9 * (1.0 - 2.0 * margin / max(ny, 2 * margin + 1))
    az = 0.9 * (1.0 - 2.0 * margin / max(nz, 2 * margin + 1))

    # trilinear up-sampling of the coarse texture, separably
    def interp_axis(n):
        pos = np.linspace(0.0, 15.0, n)
        i0 = np.minimum(np.floor(pos).astype(np.int64), 14)
        w = pos - i0
        return i0, w

    ix, wx = interp_axis(nx)
    iy, wy = interp_axis(ny)
    iz, wz = interp_axis(nz)
    tex_x = coarse[:, :, ix] * (1 - wx) + coarse[:, :, ix + 1] * wx          # (16,16,nx)
    tex_xy = tex_x[:, iy, :] * (1 - wy)[None, :, None] + tex_x[:, iy + 1, :] * wy[None, :, None]  # (16,ny,nx)

    data = np.zeros((nz, ny, nx), dtype=f32)
    X2 = (xs[None, :] / ax) ** 2
    Y2 = (ys[:, None] / ay) ** 2
    for k in range(nz):
        tex = tex_xy[iz[k]] * (1 - wz[k]) + tex_xy[iz[k] + 1] * wz[k]
        body = (X2 + Y2 + (zs[k] / az) ** 2) <= 1.0
        sl = np.where(body, 0.02 + tex, 0.0)
        for c, r in bones:
            dz2 = ((zs[k] - c[2]) / r[2]) ** 2
            if dz2 > 1.0:
                continue
            inside = (((xs[None, :] - c[0]) / r[0]) ** 2 + ((ys[:, None] - c[1]) / r[1]) ** 2 + dz2) <= 1.0
            sl = np.where(inside & body, 0.05, sl)
        data[k] = sl.astype(f32)
    sp = np.asarray(spacing, dtype=np.float64)
    origin = -0.5 * (np.array([nx, ny, nz], dtype=np.float64) - 1.0) * sp
    return Volume(data=data, spacing=tuple(sp), origin=tuple(origin), direction=np.eye(3))


def make_camera(det_size: int, full_size: int = 1536, sdd: float = 1020.0, pixel: float = 0.194) -> CameraModel:
    """Reference C-arm geometry (lib/file_formats/xregCIOSFusionDICOM.cpp:49-64) down-sampled
    to det_size x det_size with DownsampleCameraModel."""
    cam = CameraModel(coord_frame_type=kORIGIN_AT_FOCAL_PT_DET_NEG_Z).setup(sdd, full_size, full_size, pixel, pixel)
    if det_size == full_size:
        return cam
    return downsample_camera_model(cam, det_size / float(full_size))


def multi_view_cameras(det_size: int, angles_deg: Sequence[float], src_to_iso: float = 650.0) -> list:
    """Cameras of a multi-view acquisition: the C-arm of make_camera()
    rotated by each angle about the axis through the isocentre that is parallel to the volume's
    long axis in the nominal pose (the camera frame's y axis).  View 0 (angle 0) has identity
    extrinsics; all share the world frame of view 0, as in the reference's multi-view apps."""
    base = make_camera(det_size)
    iso = np.array([0.0, 0.0, -src_to_iso])
    cams = []
    for ang in angles_deg:
        E = np.eye(4)
        if ang != 0.0:
            C, Ci = np.eye(4), np.eye(4)
            C[:3, 3], Ci[:3, 3] = iso, -iso
            E = C @ rot_about_axis(1, ang) @ Ci
        cam = CameraModel(coord_frame_type=base.coord_frame_type)
        cam.setup_intrins_extrins(base.intrins, E.astype(f32), base.num_det_rows, base.num_det_cols,
                                  base.det_row_spacing, base.det_col_spacing)
        cams.append(cam)
    return cams


def rot_about_axis(axis: int, deg: float) -> np.ndarray:
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    R = np.eye(4, dtype=np.float64)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def nominal_pose(vol: Volume, src_to_iso: float = 650.0, view_rot_deg: float = 0.0) -> np.ndarray:
    """cam -> volume-physical transform: camera looks along +y of the volume (AP view),
    detector rows along -z; volume centre on the optical axis src_to_iso mm from the source.
    view_rot_deg rotates the C-arm about the volume's long (z) axis."""
    nx, ny, nz = vol.dims
    sp = np.asarray(vol.spacing, dtype=np.float64)
    centre = np.asarray(vol.origin, dtype=np.float64) + 0.5 * (np.array([nx, ny, nz]) - 1.0) * sp
    R = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], dtype=np.float64)
    T = np.eye(4, dtype=np.float64)
    T[:3, :3] = R
    T[:3, 3] = centre - R @ np.array([0.0, 0.0, -src_to_iso])
    if view_rot_deg != 0.0:
        C = np.eye(4)
        C[:3, 3] = centre
        Ci = np.eye(4)
        Ci[:3, 3] = -centre
        T = C @ rot_about_axis(2, view_rot_deg) @ Ci @ T
    return T.astype(f32)


def pose_population(vol: Volume, nominal: np.ndarray, n: int, seed: int = SEED,
                    sigma: Tuple[float, ...] = (5.0, 5.0, 5.0, 5.0, 5.0, 10.0)) -> np.ndarray:
    """n poses = rotation/translation perturbations about the volume centre,
    se(3) samples ~ N(0, diag(sigma^2)) (degrees, mm), composed as C exp(x) C^-1 nominal
    (pre * delta * post, xregIntensity2D3DRegi.cpp:1049-1071).  Returns (n, 4, 4) float32."""
    rng = np.random.default_rng(seed + 1)
    nx, ny, nz = vol.dims
    sp = np.asarray(vol.spacing, dtype=np.float64)
    centre = np.asarray(vol.origin, dtype=np.float64) + 0.5 * (np.array([nx, ny, nz]) - 1.0) * sp
    C = np.eye(4, dtype=f32)
    C[:3, 3] = centre
    Ci = np.eye(4, dtype=f32)
    Ci[:3, 3] = -centre
    sig = np.asarray(sigma, dtype=np.float64)
    sig[:3] = np.deg2rad(sig[:3])
    out = np.zeros((n, 4, 4), dtype=f32)
    for i in range(n):
        x = rng.standard_normal(6) * sig
        out[i] = (C @ exp_se3(x) @ Ci @ nominal).astype(f32)
    return out


def add_noise(img: np.ndarray, frac: float = 0.01, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed + 2)
    return (img + rng.standard_normal(img.shape) * (frac * float(img.max()))).astype(f32)


def circular_mask(rows: int, cols: int, frac: float = 0.9) -> np.ndarray:
    r = (np.arange(rows) - (rows - 1) / 2.0)[:, None]
    c = (np.arange(cols) - (cols - 1) / 2.0)[None, :]
    return ((r * r + c * c) <= (frac * min(rows, cols) / 2.0) ** 2).astype(np.uint8)


def patch_radius_for(det_size: int, full_size: int = 1536) -> int:
    """lround(41 * ds) rule of the reference app (pelvis...main.cpp:268)."""
    return int(np.floor(41.0 * det_size / full_size + 0.5))
