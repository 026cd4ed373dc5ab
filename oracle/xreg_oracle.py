"""ctypes binding of the CPU oracle (oracle/libxreg_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under xreg_b200/
imports this module.  The DRR, NCC, SSD, patch-NCC, gradient-NCC, patch gradient-NCC
and HU-conversion restatements are pinned to the reference's own source lines
(oracle/ref_pin/, tests/test_oracle_ref_slice.py); the Gaussian / Sobel filter
arithmetic (OpenCV) is PARITY UNPINNED by the reference (it has no tests for this
path) and pinned by the cv2 binding; see xreg_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libxreg_oracle.so")


class XoCam(C.Structure):
    _fields_ = [
        ("rows", C.c_uint32),
        ("cols", C.c_uint32),
        ("intrins_inv", C.c_float * 9),
        ("extrins_inv", C.c_float * 12),
        ("pinhole", C.c_float * 3),
        ("focal_len", C.c_float),
        ("frame_type", C.c_int32),
    ]


class XoPatchOpts(C.Structure):
    _fields_ = [
        ("radius", C.c_uint32),
        ("stride", C.c_uint32),
        ("compute_mean_of_patch_sims", C.c_int32),
        ("weight_patch_sims", C.c_int32),
        ("use_mask_for_weighting", C.c_int32),
        ("use_mask_for_patch_stats", C.c_int32),
        ("normalize_weights_as_prob", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc, no -march, no FMA contraction)."""
    src = os.path.join(_HERE, "xreg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.xo_interp_linear.restype = C.c_double
        _lib.xo_num_patches.restype = C.c_uint64
        _lib.xo_drr.restype = C.c_int
        _lib.xo_drr_interp.restype = C.c_int
        _lib.xo_depth.restype = C.c_int
        _lib.xo_interp_nn.restype = C.c_double
        _lib.xo_num_threads.restype = C.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def cam_struct(cam) -> XoCam:
    """cam: any object with the CameraModel fields of xreg_b200.geometry.CameraModel."""
    s = XoCam()
    s.rows = int(cam.num_det_rows)
    s.cols = int(cam.num_det_cols)
    s.intrins_inv[:] = [float(x) for x in np.asarray(cam.intrins_inv, dtype=np.float32).reshape(9)]
    s.extrins_inv[:] = [float(x) for x in np.asarray(cam.extrins_inv, dtype=np.float32)[:3, :].reshape(12)]
    s.pinhole[:] = [float(x) for x in np.asarray(cam.pinhole_pt, dtype=np.float32).reshape(3)]
    s.focal_len = float(cam.focal_len)
    s.frame_type = int(cam.coord_frame_type)
    return s


def num_threads() -> int:
    return int(lib().xo_num_threads())


def set_num_threads(n: int) -> None:
    """Use n OpenMP threads from now on, whatever OMP_NUM_THREADS said at start-up."""
    lib().xo_set_num_threads(C.c_int(int(n)))


def host_cores() -> int:
    """Cores this process may run on (affinity mask), the thread count of the timed CPU legs."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def affine_inverse(a12):
    a = _f32(a12).reshape(12)
    out = np.zeros(12, np.float32)
    lib().xo_affine_inverse(_fp(a), _fp(out))
    return out


def cam_setup_naive(focal_len, rows, cols, row_spacing, col_spacing, frame_type=1) -> XoCam:
    s = XoCam()
    lib().xo_cam_setup_naive(C.byref(s), C.c_float(focal_len), C.c_uint32(rows), C.c_uint32(cols),
                             C.c_float(row_spacing), C.c_float(col_spacing), C.c_int32(frame_type))
    return s


def cam_setup(intrins, extrins, rows, cols, row_spacing, col_spacing, frame_type=1) -> XoCam:
    s = XoCam()
    k = _f32(intrins).reshape(9)
    e = _f32(extrins).reshape(16)
    lib().xo_cam_setup(C.byref(s), _fp(k), _fp(e), C.c_uint32(rows), C.c_uint32(cols),
                       C.c_float(row_spacing), C.c_float(col_spacing), C.c_int32(frame_type))
    return s


def distribute_xforms(poses, n_cams):
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    out = np.zeros((n * n_cams, 12), np.float32)
    idx = np.zeros(n * n_cams, np.uint32)
    lib().xo_distribute_xforms(_fp(poses), C.c_uint32(n), C.c_uint32(n_cams), _fp(out),
                               idx.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out, idx


def pre_compute(buf, cam_idx, bg_projs=None, store_method=0, default_bg=0.0):
    """buf: (n_projs, rows, cols) float32, modified in place."""
    n, rows, cols = buf.shape
    cam_idx = np.ascontiguousarray(cam_idx, dtype=np.uint32)
    if bg_projs is not None:
        bgs = [_f32(b) for b in bg_projs]
        arr = (C.POINTER(C.c_float) * len(bgs))(*[_fp(b) for b in bgs])
    else:
        arr = None
    lib().xo_pre_compute(_fp(buf), C.c_uint32(n), C.c_uint32(rows), C.c_uint32(cols),
                         cam_idx.ctypes.data_as(C.POINTER(C.c_uint32)), arr, C.c_int(store_method),
                         C.c_float(default_bg))


def drr(vol, idx_to_phys, cams, poses, cam_idx=None, step_size=1.0, kernel_id=0, buf=None,
        want_info=False, n_threads=0, interp=0):
    """vol: (nz, ny, nx) float32.  cams: list of XoCam.  poses: (n, 12).

    Returns buf (n, rows, cols) [and (hit_mask, num_samples, S) if want_info].
    If buf is None a zero-initialised REPLACE buffer is used.
    """
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    a = _f32(idx_to_phys).reshape(12)
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    cam_arr = (XoCam * len(cams))(*cams)
    rows, cols = cams[0].rows, cams[0].cols
    if cam_idx is None:
        cam_idx = np.zeros(n, np.uint32)
    cam_idx = np.ascontiguousarray(cam_idx, dtype=np.uint32)
    if buf is None:
        buf = np.zeros((n, rows, cols), np.float32)
    assert buf.dtype == np.float32 and buf.flags.c_contiguous and buf.shape == (n, rows, cols)
    mask = np.zeros((n, rows, cols), np.uint8) if want_info else None
    steps = np.zeros((n, rows, cols), np.uint32) if want_info else None
    S = C.c_uint64(0)
    rc = lib().xo_drr_interp(_fp(vol), dims, _fp(a), cam_arr, C.c_uint32(len(cams)), _fp(poses),
                      cam_idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(n), C.c_float(step_size),
                      C.c_int(kernel_id), C.c_int(interp), _fp(buf),
                      _u8p(mask) if want_info else None,
                      steps.ctypes.data_as(C.POINTER(C.c_uint32)) if want_info else None,
                      C.byref(S), C.c_int(n_threads))
    if rc != 0:
        raise ValueError("xo_drr failed with code %d" % rc)
    if want_info:
        return buf, mask, steps, int(S.value)
    return buf


RAY_CAST_MAX_DEPTH = np.float32(1.0e37)   # xregRayCastInterface.h:601


def depth(vol, idx_to_phys, cams, poses, cam_idx=None, step_size=1.0, interp=0, thresh=150.0, n_backtrack=0, buf=None,
          n_threads=0):
    """RayCasterDepthCPU::compute (xo_depth).  buf (n, rows, cols): min-combined in place; None = a buffer initialised with
    kRAY_CAST_MAX_DEPTH (the class's default background).  thresh / n_backtrack default like RayCasterDepthCPU
    (xregRayCastInterface.cpp:427-428)."""
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    a = _f32(idx_to_phys).reshape(12)
    poses = _f32(poses).reshape(-1, 12)
    n = poses.shape[0]
    cam_arr = (XoCam * len(cams))(*cams)
    rows, cols = cams[0].rows, cams[0].cols
    if cam_idx is None:
        cam_idx = np.zeros(n, np.uint32)
    cam_idx = np.ascontiguousarray(cam_idx, dtype=np.uint32)
    if buf is None:
        buf = np.full((n, rows, cols), RAY_CAST_MAX_DEPTH, np.float32)
    assert buf.dtype == np.float32 and buf.flags.c_contiguous and buf.shape == (n, rows, cols)
    rc = lib().xo_depth(_fp(vol), dims, _fp(a), cam_arr, C.c_uint32(len(cams)), _fp(poses),
                        cam_idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(n), C.c_float(step_size), C.c_int(interp),
                        C.c_float(thresh), C.c_uint32(n_backtrack), _fp(buf), C.c_int(n_threads))
    if rc != 0:
        raise ValueError("xo_depth failed with code %d" % rc)
    return buf


def itk_discrete_gaussian_2d(img, variance):
    """itk::DiscreteGaussianImageFilter (UseImageSpacing off, max error 0.01, max kernel width 32) restated; PARITY
    UNPINNED (ITK is absent)."""
    img = _f32(img)
    out = np.empty_like(img)
    lib().xo_itk_discrete_gaussian_2d(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_double(variance), _fp(out))
    return out


def downsample_image(img, factor, sigma=-1.0):
    """DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, B-spline default): ITK's Gaussian + cubic B-spline resampling
    restated; PARITY UNPINNED."""
    img = _f32(img)
    orows, ocols = C.c_uint32(0), C.c_uint32(0)
    lib().xo_downsample_size(C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_double(factor), C.byref(orows), C.byref(ocols))
    out = np.empty((orows.value, ocols.value), np.float32)
    lib().xo_downsample_image(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_double(factor), C.c_double(sigma), _fp(out))
    return out


def log_remap(img, normalize_zero_one=False, use_max_intensity_as_I0=True, I0=1.0, smoothed=None):
    """ImageIntensLogTransFilter (lib/image/xregImageIntensLogTrans.cpp:55-144).  Returns (out, I0 used)."""
    img = _f32(img)
    out = np.empty_like(img)
    i0 = C.c_float(0)
    sm = None if smoothed is None else _f32(smoothed)
    lib().xo_log_remap(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_int(1 if normalize_zero_one else 0),
                       C.c_int(1 if use_max_intensity_as_I0 else 0), C.c_float(I0), None if sm is None else _fp(sm), _fp(out),
                       C.byref(i0))
    return out, np.float32(i0.value)


def interp_linear(vol, x):
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    xx = _f32(x).reshape(3)
    return float(lib().xo_interp_linear(_fp(vol), dims, _fp(xx)))


def interp_nn(vol, x):
    vol = _f32(vol)
    nz, ny, nx = vol.shape
    dims = (C.c_uint64 * 3)(nx, ny, nz)
    xx = _f32(x).reshape(3)
    return float(lib().xo_interp_nn(_fp(vol), dims, _fp(xx)))


def ncc(fixed, mov, mask=None, n_threads=0, inplace=False):
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    if not inplace:
        mov = mov.copy()
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    lib().xo_ncc(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols), _fp(mov),
                 C.c_uint32(mov.shape[0]), _fp(sims), C.c_int(n_threads))
    return sims


def hu_to_lin_att(hu, hu_lower=-1000.0):
    """HUToLinAtt (lib/image/xregHUToLinAtt.cpp:45-69)."""
    hu = _f32(hu)
    out = np.zeros_like(hu)
    lib().xo_hu_to_lin_att(_fp(hu), _fp(out), C.c_uint64(hu.size), C.c_float(hu_lower))
    return out


def ssd(fixed, mov, mask=None, n_threads=0):
    """ImgSimMetric2DSSDCPU: sum((fixed - mov)^2) / num_pixels, images zeroed outside the mask."""
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols).copy()
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    lib().xo_ssd(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols), _fp(mov),
                 C.c_uint32(mov.shape[0]), _fp(sims), C.c_int(n_threads))
    return sims


def gauss_kernel(width):
    cf = np.zeros(width, np.float32)
    if lib().xo_gauss_kernel(C.c_int(width), _fp(cf)) != 0:
        raise ValueError("bad Gaussian width")
    return cf


def gauss_blur(img, width):
    img = _f32(img)
    out = np.zeros_like(img)
    lib().xo_gauss_blur(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_int(width), _fp(out))
    return out


def sobel(img):
    img = _f32(img)
    gx = np.zeros_like(img)
    gy = np.zeros_like(img)
    lib().xo_sobel(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), _fp(gx), _fp(gy))
    return gx, gy


def grad_imgs(img, gauss_width=5):
    img = _f32(img)
    gx = np.zeros_like(img)
    gy = np.zeros_like(img)
    lib().xo_grad_imgs(_fp(img), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]), C.c_int(gauss_width),
                       _fp(gx), _fp(gy))
    return gx, gy


def grad_ncc(fixed, mov, mask=None, gauss_width=5, n_threads=0):
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    sims = np.zeros(mov.shape[0], np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    lib().xo_grad_ncc(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
                      C.c_int(gauss_width), _fp(mov), C.c_uint32(mov.shape[0]), _fp(sims), C.c_int(n_threads))
    return sims


def patch_opts(radius=5, stride=1, compute_mean=False, weight_sims=True, mask_weighting=True,
               mask_stats=False, normalize=True) -> XoPatchOpts:
    return XoPatchOpts(radius, stride, int(compute_mean), int(weight_sims), int(mask_weighting),
                       int(mask_stats), int(normalize))


def num_patches(rows, cols, radius, stride=1):
    return int(lib().xo_num_patches(C.c_uint32(rows), C.c_uint32(cols), C.c_uint32(radius), C.c_uint32(stride)))


def patch_weights(rows, cols, opts, mask=None, wgt_img=None):
    n = num_patches(rows, cols, opts.radius, opts.stride)
    w = np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    wi = _f32(wgt_img) if wgt_img is not None else None
    lib().xo_patch_weights(C.c_uint32(rows), C.c_uint32(cols), C.byref(opts), _u8p(m) if m is not None else None,
                           _fp(wi) if wi is not None else None, _fp(w))
    return w


def patch_ncc(fixed, mov, opts, mask=None, weights=None, want_patch_sims=False, n_threads=0):
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    n = mov.shape[0]
    sims = np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    w = _f32(weights) if weights is not None else None
    ps = np.zeros((n, num_patches(rows, cols, opts.radius, opts.stride)), np.float32) if want_patch_sims else None
    lib().xo_patch_ncc(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
                       C.byref(opts), _fp(w) if w is not None else None, _fp(mov), C.c_uint32(n), _fp(sims),
                       _fp(ps) if ps is not None else None, C.c_int(n_threads))
    return (sims, ps) if want_patch_sims else sims


def patch_grad_ncc(fixed, mov, opts, mask=None, weights=None, gauss_width=5, n_threads=0):
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    n = mov.shape[0]
    sims = np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    w = _f32(weights) if weights is not None else None
    lib().xo_patch_grad_ncc(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
                            C.c_int(gauss_width), C.byref(opts), _fp(w) if w is not None else None, _fp(mov),
                            C.c_uint32(n), _fp(sims), C.c_int(n_threads))
    return sims


def patch_ncc_subset(fixed, mov, opts, subset, mask=None, weights=None, gauss_width=None, n_threads=0):
    """Patch NCC (gauss_width None) or patch gradient-NCC over the patch subset `subset` (global patch indices, local
    order = list order, repeats allowed): set_patches_to_use / random patches of the reference."""
    fixed = _f32(fixed)
    rows, cols = fixed.shape
    mov = _f32(mov).reshape(-1, rows, cols)
    n = mov.shape[0]
    sims = np.zeros(n, np.float32)
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    w = _f32(weights) if weights is not None else None
    sub = np.ascontiguousarray(subset, dtype=np.uint64)
    sp = sub.ctypes.data_as(C.POINTER(C.c_uint64))
    if gauss_width is None:
        lib().xo_patch_ncc_subset(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
                                  C.byref(opts), _fp(w) if w is not None else None, sp, C.c_uint64(sub.size), _fp(mov),
                                  C.c_uint32(n), _fp(sims), C.c_int(n_threads))
    else:
        lib().xo_patch_grad_ncc_subset(_fp(fixed), _u8p(m) if m is not None else None, C.c_uint32(rows), C.c_uint32(cols),
                                       C.c_int(gauss_width), C.byref(opts), _fp(w) if w is not None else None, sp,
                                       C.c_uint64(sub.size), _fp(mov), C.c_uint32(n), _fp(sims), C.c_int(n_threads))
    return sims


def combine_mean(view_sims):
    v = _f32(view_sims)
    out = np.zeros(v.shape[1], np.float32)
    lib().xo_combine_mean(_fp(v), C.c_uint32(v.shape[0]), C.c_uint32(v.shape[1]), _fp(out))
    return out
