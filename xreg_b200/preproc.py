"""Projection pre-processing behind the same boundary: the log remap xReg applies to every fluoroscopic
image before a registration (ProjPreProc, lib/image/xregProjPreProc.cpp:63-84 -> ImageIntensLogTransFilter,
lib/image/xregImageIntensLogTrans.{h,cpp}), on the device (xrc_log_remap)."""
import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from .ray_caster import Context

f32 = np.float32


def log_remap(ctx: Context, img: np.ndarray, normalize_zero_one: bool = False, use_max_intensity_as_I0: bool = True,
              I0: float = 1.0) -> Tuple[np.ndarray, np.float32]:
    """ImageIntensLogTransFilter: SetNormalizeZeroOne / SetUseMaxIntensityAsI0 / SetI0, Update().  Returns (the remapped
    image, the I0 that was used)."""
    a = np.ascontiguousarray(img, dtype=f32)
    if a.ndim != 2 or a.size == 0:
        raise _lib.XregError("log_remap: a non-empty 2-D image is expected")
    out = np.empty_like(a)
    i0 = C.c_float(0)
    FP = C.POINTER(C.c_float)
    _lib.check(_lib.load().xrc_log_remap(ctx.handle, a.ctypes.data_as(FP), a.shape[0], a.shape[1],
                                         1 if normalize_zero_one else 0, 1 if use_max_intensity_as_I0 else 0, float(I0),
                                         out.ctypes.data_as(FP), C.byref(i0)))
    return out, f32(i0.value)


def downsample_image(ctx: Context, img: np.ndarray, factor: float, sigma: float = -1.0) -> np.ndarray:
    """DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, cubic B-spline): Gaussian smoothing (sigma < 0: the default
    0.5 / factor) + resampling, on the device (xrc_downsample_image)."""
    a = np.ascontiguousarray(img, dtype=f32)
    if a.ndim != 2 or a.size == 0:
        raise _lib.XregError("downsample_image: a non-empty 2-D image is expected")
    lib = _lib.load()
    orows, ocols = C.c_uint32(0), C.c_uint32(0)
    _lib.check(lib.xrc_downsample_size(a.shape[0], a.shape[1], float(factor), C.byref(orows), C.byref(ocols)))
    out = np.empty((orows.value, ocols.value), dtype=f32)
    FP = C.POINTER(C.c_float)
    _lib.check(lib.xrc_downsample_image(ctx.handle, a.ctypes.data_as(FP), a.shape[0], a.shape[1], float(factor), float(sigma),
                                        out.ctypes.data_as(FP)))
    return out


def downsample_proj_data(ctx: Context, img: np.ndarray, cam, ds_factor: float, force_even_dims: bool = False):
    """DownsampleProjData (lib/image/xregProjData.cpp:40-99) for one projection: the camera model through
    DownsampleCameraModel, the image through DownsampleImage, cropped to even dimensions from index (0, 0) when asked
    (:52-83).  Returns (image, camera).  (Landmarks are scaled by ds_factor on the caller's side, :90-94.)"""
    from .geometry import downsample_camera_model

    dcam = downsample_camera_model(cam, ds_factor, force_even_dims)
    dimg = downsample_image(ctx, img, ds_factor)
    if force_even_dims:
        r, c = dimg.shape
        dimg = np.ascontiguousarray(dimg[: r - (r % 2), : c - (c % 2)])
    if dimg.shape != (dcam.num_det_rows, dcam.num_det_cols):
        raise _lib.XregError("downsample_proj_data: image %s and camera (%d, %d) disagree (xregProjData.cpp:86-87)"
                             % (dimg.shape, dcam.num_det_rows, dcam.num_det_cols))
    return dimg, dcam
