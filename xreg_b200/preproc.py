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
