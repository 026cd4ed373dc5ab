"""TEST INFRASTRUCTURE.  ctypes binding of oracle/_ref/libxreg_refslice.so: the reference's own RayRectIntersect,
CameraModel::ind_pt_to_phys_det_pt and ComputeLineInts<Kernel>, compiled from /root/reference by build_ref_slice.py
over the stand-in types of ref_pin_prelude.h.  Used only by tests/test_oracle_ref_slice.py to pin the oracle."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref_slice
from ..xreg_oracle import XoCam, _f32, _fp

_lib = None


def available() -> bool:
    return os.path.exists(build_ref_slice.LIB) or os.path.isdir(build_ref_slice.REF)


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir(build_ref_slice.REF):   # rebuild where the reference exists (cheap); elsewhere use the shipped .so
            build_ref_slice.build()
        _lib = C.CDLL(build_ref_slice.LIB)
        _lib.xref_ray_rect_intersect.restype = C.c_int
        _lib.xref_compute_line_ints.restype = C.c_int
        _lib.xref_compute_line_ints_interp.restype = C.c_int
        _lib.xref_compute_depth.restype = C.c_int
    return _lib


def ray_rect_intersect(mn, mx, p, d, limit_to_segment=True):
    t0, t1 = C.c_float(0), C.c_float(0)
    a, b, c, e = (_f32(v).reshape(3) for v in (mn, mx, p, d))
    hit = lib().xref_ray_rect_intersect(_fp(a), _fp(b), _fp(c), _fp(e), C.c_int(1 if limit_to_segment else 0),
                                        C.byref(t0), C.byref(t1))
    return bool(hit), np.float32(t0.value), np.float32(t1.value)


def ind_pt_to_phys_det_pt(cam: XoCam, col: float, row: float) -> np.ndarray:
    out = np.zeros(3, np.float32)
    lib().xref_ind_pt_to_phys_det_pt(C.byref(cam), C.c_float(col), C.c_float(row), _fp(out))
    return out


def compute_line_ints(vol, phys_to_idx, cams, poses, cam_idx=None, step_size=1.0, kernel_id=0, buf=None, interp=0):
    """ComputeLineInts<Accum|Max kernel> over the whole projection range (serial).  vol (nz, ny, nx) f32; phys_to_idx the
    already inverted 3x4 (row-major); cams list of XoCam; poses (n, 12).  buf: initialised projection buffer or None
    (zeros: the REPLACE store without background)."""
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    cam_arr = (XoCam * len(cams))(*cams)
    rows, cols = cams[0].rows, cams[0].cols
    ci = np.ascontiguousarray(np.zeros(n, np.uint32) if cam_idx is None else cam_idx, dtype=np.uint32)
    if buf is None:
        buf = np.zeros((n, rows, cols), np.float32)
    a = _f32(phys_to_idx).reshape(12)
    rc = lib().xref_compute_line_ints_interp(_fp(vol), dims, _fp(a), cam_arr, C.c_uint32(len(cams)), _fp(poses),
                                             ci.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(n), C.c_float(step_size),
                                             C.c_int(kernel_id), C.c_int(interp), _fp(buf))
    if rc != 0:
        raise ValueError("xref_compute_line_ints failed")
    return buf


def compute_depth(vol, phys_to_idx, cams, poses, cam_idx=None, step_size=1.0, interp=0, thresh=150.0, n_backtrack=0, buf=None):
    """RayCastDepthFn over the whole projection range (serial); buf initialised by the caller or with kRAY_CAST_MAX_DEPTH."""
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    cam_arr = (XoCam * len(cams))(*cams)
    rows, cols = cams[0].rows, cams[0].cols
    ci = np.ascontiguousarray(np.zeros(n, np.uint32) if cam_idx is None else cam_idx, dtype=np.uint32)
    if buf is None:
        buf = np.full((n, rows, cols), np.float32(1.0e37), np.float32)
    a = _f32(phys_to_idx).reshape(12)
    rc = lib().xref_compute_depth(_fp(vol), dims, _fp(a), cam_arr, C.c_uint32(len(cams)), _fp(poses),
                                  ci.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(n), C.c_float(step_size),
                                  C.c_int(interp), C.c_float(thresh), C.c_uint32(n_backtrack), _fp(buf))
    if rc != 0:
        raise ValueError("xref_compute_depth failed")
    return buf


# ---- the reference's patch-NCC class (oracle/_ref/libxreg_refslice_metric.so) ----------------------------------------
_mlib = None


def metric_lib():
    global _mlib
    if _mlib is None:
        lib()   # (re)builds both units where the reference exists
        _mlib = C.CDLL(build_ref_slice.METRIC_LIB)
        _mlib.xref_patch_ncc.restype = C.c_int
    return _mlib


def patch_mean_std(img, mask, r0, c0, d, use_mask_for_stats):
    """detail::ComputePatchMeanStdDev on the d x d patch at (r0, c0): (mean, clamped std dev, pixels counted)."""
    img = _f32(img)
    rows, cols = img.shape
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    mean, sd, n = C.c_float(0), C.c_float(0), C.c_uint64(0)
    metric_lib().xref_patch_mean_std(_fp(img), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None,
                                     C.c_uint32(rows), C.c_uint32(cols), C.c_uint32(r0), C.c_uint32(c0), C.c_uint32(d),
                                     C.c_int(1 if use_mask_for_stats else 0), C.byref(mean), C.byref(sd), C.byref(n))
    return np.float32(mean.value), np.float32(sd.value), int(n.value)


def patch_ncc_subset(fixed, mov, opts, subset, mask=None, wgt_img=None):
    """ImgSimMetric2DPatchNCCCPU with set_patches_to_use(subset) before allocate_resources() / compute()."""
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    n = mov.shape[0]
    sims = np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    wi = _f32(wgt_img) if wgt_img is not None else None
    sub = np.ascontiguousarray(subset, dtype=np.uint64)
    fn = metric_lib().xref_patch_ncc
    fn(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
       C.byref(opts), _fp(wi) if wi is not None else None, _fp(mov), C.c_uint32(n), _fp(sims), None, None,
       sub.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_uint64(sub.size))
    return sims


def patch_ncc(fixed, mov, opts, mask=None, wgt_img=None, want_patch_sims=False):
    """ImgSimMetric2DPatchNCCCPU: set_fixed_image / set_mask / set_wgt_img / patch parameters, allocate_resources(),
    compute().  Returns (sims, patch weights[, per-patch values (n, num_patches)])."""
    from ..xreg_oracle import num_patches

    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    n = mov.shape[0]
    np_ = num_patches(rows, cols, opts.radius, opts.stride)
    sims = np.zeros(n, np.float32)
    w = np.zeros(np_, np.float32)
    ps = np.zeros((n, np_), np.float32) if want_patch_sims else None
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    wi = _f32(wgt_img) if wgt_img is not None else None
    got = metric_lib().xref_patch_ncc(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None,
                                      C.c_uint32(rows), C.c_uint32(cols), C.byref(opts), _fp(wi) if wi is not None else None,
                                      _fp(mov), C.c_uint32(n), _fp(sims), _fp(w), _fp(ps) if ps is not None else None,
                                      None, C.c_uint64(0))
    assert got == np_, "patch grid size differs: reference %d, oracle %d" % (got, np_)
    return (sims, w, ps) if want_patch_sims else (sims, w)


def hu_to_lin_att(hu, hu_lower=-1000.0):
    """HUToLinAtt(hu_vol, hu_lower) through the reference's HUToLinAttFilter::GenerateData."""
    lib()
    hl = C.CDLL(build_ref_slice.HU_LIB)
    hu = _f32(hu)
    out = np.zeros_like(hu)
    hl.xref_hu_to_lin_att(_fp(hu), _fp(out), C.c_uint64(hu.size), C.c_float(hu_lower))
    return out


def ncc(fixed, mov, mask=None):
    """ImgSimMetric2DNCCCPU: set_fixed_image / set_mask, allocate_resources(), compute() -> sim_vals."""
    lib()
    nl = C.CDLL(build_ref_slice.NCC_LIB)
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    nl.xref_ncc(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None, C.c_uint32(rows),
                C.c_uint32(cols), _fp(mov), C.c_uint32(mov.shape[0]), _fp(sims))
    return sims


def distribute_xforms(poses, n_cams):
    """RayCaster::distribute_xforms_among_cam_models: (n, 12) poses -> (n_cams * n, 12) poses + camera indices."""
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    out = np.zeros((n * n_cams, 12), np.float32)
    idx = np.zeros(n * n_cams, np.uint32)
    lib().xref_distribute_xforms(_fp(poses), C.c_uint32(n), C.c_uint32(n_cams), _fp(out),
                                 idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out, idx


def pre_compute(buf, cam_idx, n_cams, bg_projs=None, store_method=0, default_bg=0.0):
    """RayCasterCPU::pre_compute on buf (n_projs, rows, cols), in place."""
    n, rows, cols = buf.shape
    ci = np.ascontiguousarray(cam_idx, dtype=np.uint32)
    arr = None
    if bg_projs is not None:
        bgs = [_f32(b) for b in bg_projs]
        arr = (C.POINTER(C.c_float) * len(bgs))(*[_fp(b) for b in bgs])
    lib().xref_pre_compute(_fp(buf), C.c_uint32(n), C.c_uint32(rows), C.c_uint32(cols),
                           ci.ctypes.data_as(C.POINTER(C.c_uint32)), arr, C.c_uint32(n_cams), C.c_int(store_method),
                           C.c_float(default_bg))


def combine(view_sims, mean=True):
    """ImgSimMetric2DCombineMean (mean=True) / ImgSimMetric2DCombineAddition over per-view similarity values (views, poses)."""
    lib()
    nl = C.CDLL(build_ref_slice.NCC_LIB)
    v = _f32(view_sims)
    out = np.zeros(v.shape[1], np.float32)
    nl.xref_combine(_fp(v), C.c_uint32(v.shape[0]), C.c_uint32(v.shape[1]), C.c_int(1 if mean else 0), _fp(out))
    return out


# ---- the gradient-metric classes (oracle/_ref/libxreg_refslice_grad.so); cv::GaussianBlur / cv::Sobel are call-outs --
_GAUSS_T = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float))
_SOBEL_T = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float))
_glib = None
_cv_keep = None


def grad_lib():
    global _glib
    if _glib is None:
        lib()
        _glib = C.CDLL(build_ref_slice.GRAD_LIB)
        _glib.xref_patch_grad_ncc.restype = C.c_int
    return _glib


def set_filters(gauss, sobel):
    """Install the Gaussian / Sobel the reference classes call: gauss(img (rows, cols) f32, ksize) -> img,
    sobel(img, dx, dy) -> img (3x3, scale 1, default border)."""
    global _cv_keep

    def g(src, rows, cols, ksize, dst):
        a = np.ctypeslib.as_array(src, shape=(rows, cols)).copy()
        np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = np.asarray(gauss(a, ksize), dtype=np.float32)

    def s(src, rows, cols, dx, dy, dst):
        a = np.ctypeslib.as_array(src, shape=(rows, cols)).copy()
        np.ctypeslib.as_array(dst, shape=(rows, cols))[:] = np.asarray(sobel(a, dx, dy), dtype=np.float32)

    _cv_keep = (_GAUSS_T(g), _SOBEL_T(s))
    grad_lib().xref_set_cv(_cv_keep[0], _cv_keep[1])


def grad_ncc(fixed, mov, mask=None, gauss_width=5):
    """ImgSimMetric2DGradNCCCPU (set_filters first)."""
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    grad_lib().xref_grad_ncc(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None,
                             C.c_uint32(rows), C.c_uint32(cols), C.c_int(gauss_width), _fp(mov),
                             C.c_uint32(mov.shape[0]), _fp(sims))
    return sims


def patch_grad_ncc(fixed, mov, opts, mask=None, gauss_width=5):
    """ImgSimMetric2DPatchGradNCCCPU (set_filters first)."""
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    grad_lib().xref_patch_grad_ncc(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None,
                                   C.c_uint32(rows), C.c_uint32(cols), C.c_int(gauss_width), C.byref(opts), _fp(mov),
                                   C.c_uint32(mov.shape[0]), _fp(sims))
    return sims


def ssd(fixed, mov, mask=None):
    """ImgSimMetric2DSSDCPU: allocate_resources(), compute() -> sim_vals."""
    lib()
    nl = C.CDLL(build_ref_slice.NCC_LIB)
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    nl.xref_ssd(_fp(fixed), m.ctypes.data_as(C.POINTER(C.c_uint8)) if m is not None else None, C.c_uint32(rows),
                C.c_uint32(cols), _fp(mov), C.c_uint32(mov.shape[0]), _fp(sims))
    return sims


def exp_se3(x):
    """ExpSE3(Pt6) of the reference (lib/transforms/xregRigidUtils.cpp:40-85 over xregRotUtils.cpp): row-major 3x4."""
    lib()
    sl = C.CDLL(build_ref_slice.SE3_LIB)
    xx = _f32(x).reshape(6)
    out = np.zeros(12, np.float32)
    sl.xref_exp_se3(_fp(xx), _fp(out))
    return out


# ---- camera set-up (oracle/_ref/libxreg_refslice_cam.so) ---------------------------------------------------------------
def _cam_lib():
    lib()
    return C.CDLL(build_ref_slice.CAM_LIB)


def cam_setup_naive(focal_len, rows, cols, row_spacing, col_spacing, frame_type=1) -> XoCam:
    """CameraModel::setup(focal_len, nr, nc, rs, cs) of the reference (MakeNaiveIntrins inside)."""
    s = XoCam()
    _cam_lib().xref_cam_setup_naive(C.byref(s), C.c_float(focal_len), C.c_uint32(rows), C.c_uint32(cols),
                                    C.c_float(row_spacing), C.c_float(col_spacing), C.c_int32(frame_type))
    return s


def cam_setup(intrins, extrins, rows, cols, row_spacing, col_spacing, frame_type=1) -> XoCam:
    """CameraModel::setup(intrins, extrins, nr, nc, rs, cs) of the reference (FocalLenFromIntrins, SE3Inv inside)."""
    s = XoCam()
    k, e = _f32(intrins).reshape(9), _f32(extrins).reshape(16)
    _cam_lib().xref_cam_setup(C.byref(s), _fp(k), _fp(e), C.c_uint32(rows), C.c_uint32(cols), C.c_float(row_spacing),
                              C.c_float(col_spacing), C.c_int32(frame_type))
    return s


def cam_downsample(intrins, extrins, rows, cols, row_spacing, col_spacing, frame_type, ds_factor, force_even_dims=False):
    """DownsampleCameraModel of the camera setup(intrins, extrins, ...) makes: (XoCam, intrins (3,3), (row, col) spacing)."""
    s = XoCam()
    k, e = _f32(intrins).reshape(9), _f32(extrins).reshape(16)
    ko, sp = np.zeros(9, np.float32), np.zeros(2, np.float32)
    _cam_lib().xref_cam_downsample(C.byref(s), _fp(ko), _fp(sp), _fp(k), _fp(e), C.c_uint32(rows), C.c_uint32(cols),
                                   C.c_float(row_spacing), C.c_float(col_spacing), C.c_int32(frame_type),
                                   C.c_float(ds_factor), C.c_int(1 if force_even_dims else 0))
    return s, ko.reshape(3, 3), sp


def itk_volume_geometry(dims, origin, spacing, direction):
    """ITKImageIndexBoundsAsEigen + ITKImagePhysicalPointTransformsAsEigen of the reference for an image with these meta
    data (nx, ny, nz; doubles): (aabb_min (3,), aabb_max (3,), idx_to_phys (12,) row-major 3x4)."""
    lib()
    il = C.CDLL(build_ref_slice.ITK_LIB)
    d = (C.c_uint64 * 3)(*[int(v) for v in dims])
    o = np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
    sp = np.ascontiguousarray(spacing, dtype=np.float64).reshape(3)
    dr = np.ascontiguousarray(direction, dtype=np.float64).reshape(9)
    mn, mx, a = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(12, np.float32)
    DP = C.POINTER(C.c_double)
    il.xref_itk_volume_geometry(d, o.ctypes.data_as(DP), sp.ctypes.data_as(DP), dr.ctypes.data_as(DP), _fp(mn), _fp(mx), _fp(a))
    return mn, mx, a


# ---- the reference's log remap (oracle/_ref/libxreg_refslice_log.so) ------------------------------------------------------
_llib = None
_gauss_keep = []


def log_lib():
    global _llib
    if _llib is None:
        lib()   # (re)builds every unit where the reference exists
        _llib = C.CDLL(build_ref_slice.LOG_LIB)
    return _llib


def log_remap(img, normalize_zero_one=False, use_max_intensity_as_I0=True, I0=1.0, gaussian=None):
    """ImageIntensLogTransFilter::GenerateData, the reference's lines; `gaussian(img2d, variance) -> img2d` is installed as
    the itk::DiscreteGaussianImageFilter call-out (ITK itself is absent)."""
    img = _f32(img)
    rows, cols = img.shape
    GFN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_uint, C.c_uint, C.c_double, C.POINTER(C.c_float))

    def cb(pin, r, c, var, pout):
        a = np.ctypeslib.as_array(pin, shape=(r, c)).copy()
        o = np.ascontiguousarray(gaussian(a, var), dtype=np.float32)
        np.ctypeslib.as_array(pout, shape=(r, c))[:] = o

    fn = GFN(cb)
    _gauss_keep.append(fn)
    log_lib().xref_log_remap_set_gaussian(fn)
    out = np.empty_like(img)
    log_lib().xref_log_remap(_fp(img), C.c_uint32(rows), C.c_uint32(cols), C.c_int(1 if normalize_zero_one else 0),
                             C.c_int(1 if use_max_intensity_as_I0 else 0), C.c_float(I0), _fp(out))
    return out
