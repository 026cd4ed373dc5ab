"""ImgSimMetric2D*CUDA: host-side mirrors of xreg::ImgSimMetric2D and its NCC /
Grad-NCC / Patch-NCC / Patch-Grad-NCC implementations
(lib/regi/sim_metrics_2d/xregImgSimMetric2D.h:42-156 and the *CPU / *OCL classes)
over the C ABI.  The patch-grid / weight logic of ImgSimMetric2DPatchCommon stays
on the host exactly like in the reference and is handed down as arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check
from .geometry import f32
from .ray_caster import Context, RayCasterLineIntCUDA


class ImgSimMetric2D:
    """Common part of the interface (xregImgSimMetric2D.h:42-156)."""

    KIND = _lib.SM_NCC

    def __init__(self, ctx: Context):
        self._lib = _lib.load()
        self.ctx = ctx
        h = C.c_void_p()
        check(self._lib.xrc_sm_create(ctx.handle, self.KIND, C.byref(h)))
        self.handle = h
        self._fixed: Optional[np.ndarray] = None
        self._mask: Optional[np.ndarray] = None
        self._num_mov_imgs = 0
        self._sim_vals = np.zeros(0, dtype=f32)
        self._allocated = False
        self._rc: Optional[RayCasterLineIntCUDA] = None
        self._host_buf: Optional[np.ndarray] = None

    # -- set-up ---------------------------------------------------------------
    def set_fixed_image(self, fixed_img: np.ndarray) -> None:
        self._fixed = np.ascontiguousarray(fixed_img, dtype=f32)
        r, c = self._fixed.shape
        check(self._lib.xrc_sm_set_fixed(self.handle, self._fixed.ctypes.data_as(C.POINTER(C.c_float)), r, c))

    def fixed_image(self) -> Optional[np.ndarray]:
        return self._fixed

    def set_num_moving_images(self, n: int) -> None:
        self._num_mov_imgs = int(n)
        if self._allocated:
            check(self._lib.xrc_sm_set_num_imgs(self.handle, int(n)))
            self._sim_vals = np.zeros(int(n), dtype=f32)

    def num_moving_images(self) -> int:
        return self._num_mov_imgs

    def set_mov_imgs_buf_from_ray_caster(self, ray_caster: RayCasterLineIntCUDA, proj_offset: int = 0) -> None:
        check(self._lib.xrc_sm_bind_ray_caster(self.handle, ray_caster.handle, int(proj_offset)))
        self._rc = ray_caster
        self._host_buf = None

    def set_mov_imgs_host_buf(self, mov_imgs_buf: np.ndarray, proj_offset: int = 0) -> None:
        if mov_imgs_buf.dtype != f32 or not mov_imgs_buf.flags.c_contiguous:
            raise _lib.XregError("set_mov_imgs_host_buf: need a C-contiguous float32 buffer")
        self._host_buf = mov_imgs_buf  # caller-owned, must outlive the metric (as in the reference)
        check(self._lib.xrc_sm_bind_host(self.handle, mov_imgs_buf.ctypes.data_as(C.POINTER(C.c_float)), int(proj_offset)))
        self._rc = None

    def set_mov_imgs_device_buf(self, dev_ptr: int, proj_offset: int = 0) -> None:
        check(self._lib.xrc_sm_bind_device(self.handle, C.c_void_p(dev_ptr), int(proj_offset)))
        self._rc = None
        self._host_buf = None

    def set_mask(self, mask: Optional[np.ndarray]) -> None:
        if mask is None:
            self._mask = None
            check(self._lib.xrc_sm_set_mask(self.handle, None))
        else:
            self._mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if self._fixed is None or self._mask.shape != self._fixed.shape:
                raise _lib.XregError("set_mask: mask must match the fixed image")
            check(self._lib.xrc_sm_set_mask(self.handle, self._mask.ctypes.data_as(C.POINTER(C.c_uint8))))
        self._mask_changed()

    def mask(self) -> Optional[np.ndarray]:
        return self._mask

    def _mask_changed(self) -> None:
        pass

    def _pre_allocate(self) -> None:
        pass

    def allocate_resources(self) -> None:
        if self._num_mov_imgs <= 0:
            raise _lib.XregError("allocate_resources: set_num_moving_images first")
        self._pre_allocate()
        check(self._lib.xrc_sm_allocate(self.handle, self._num_mov_imgs))
        self._sim_vals = np.zeros(self._num_mov_imgs, dtype=f32)
        self._allocated = True

    # -- compute --------------------------------------------------------------
    def compute(self) -> None:
        """Blocking like the reference: similarity values are valid on return."""
        if not self._allocated:
            raise _lib.XregError("compute: resources not allocated")
        self._pre_compute()
        check(self._lib.xrc_sm_compute(self.handle))
        if self._num_mov_imgs:
            check(self._lib.xrc_sm_read_sims(self.handle, self._sim_vals.ctypes.data_as(C.POINTER(C.c_float)),
                                             self._num_mov_imgs))

    def compute_async(self) -> None:
        self._pre_compute()
        check(self._lib.xrc_sm_compute(self.handle))

    def _pre_compute(self) -> None:
        """Hook run before every compute (the patch metrics draw their random patch subset here)."""

    def sim_val(self, mov_img_idx: int) -> float:
        return float(self._sim_vals[mov_img_idx])

    def sim_vals(self) -> np.ndarray:
        return self._sim_vals

    def device_sims(self) -> int:
        p = C.c_void_p()
        check(self._lib.xrc_sm_device_sims(self.handle, C.byref(p)))
        return int(p.value)

    def close(self) -> None:
        if self.handle:
            self._lib.xrc_sm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ImgSimMetric2DGradImgParamInterface:
    """xregImgSimMetric2DGradImgParamInterface.h:31-39 ("radius" is the kernel width)."""

    _smooth_img_kernel_rad = 5

    def smooth_img_before_sobel_kernel_radius(self) -> int:
        return self._smooth_img_kernel_rad

    def set_smooth_img_before_sobel_kernel_radius(self, r: int) -> None:
        self._smooth_img_kernel_rad = int(r)
        check(self._lib.xrc_sm_set_grad_params(self.handle, int(r)))

    def read_grads(self, img: int):
        r, c = self._fixed.shape
        gx = np.empty((r, c), dtype=f32)
        gy = np.empty((r, c), dtype=f32)
        check(self._lib.xrc_sm_read_grads(self.handle, int(img), gx.ctypes.data_as(C.POINTER(C.c_float)),
                                          gy.ctypes.data_as(C.POINTER(C.c_float))))
        return gx, gy


class ImgSimMetric2DPatchCommon:
    """Patch grid / weights / options (xregImgSimMetric2DPatchCommon.{h,cpp})."""

    def _init_patch_common(self) -> None:
        self._patch_radius = 5
        self._patch_stride = 1
        self._compute_mean_of_patch_sims = False
        self._weight_patch_sims_in_combine = True
        self._use_mask_for_weighting = True
        self._use_mask_for_patch_stats = False
        self._normalize_weights_as_prob = True
        self._wgt_img: Optional[np.ndarray] = None
        self._weights: Optional[np.ndarray] = None
        # patch subsets (xregImgSimMetric2DPatchCommon.h:100-123, .cpp:231-241, 413-493)
        self._choose_rand_patches = False
        self._num_rand_patches = 100
        self._rand_patch_min_pixels_sep = -1.0
        self._patch_inds_to_use: Optional[np.ndarray] = None
        self._do_not_update_patch_inds_to_use = False
        self._rng = np.random.default_rng()

    def patch_radius(self) -> int:
        return self._patch_radius

    def set_patch_radius(self, r: int) -> None:
        self._patch_radius = int(r)

    def patch_stride(self) -> int:
        return self._patch_stride

    def set_patch_stride(self, s: int) -> None:
        self._patch_stride = int(s)

    def set_compute_mean_of_patch_sims(self, b: bool) -> None:
        self._compute_mean_of_patch_sims = bool(b)

    def compute_mean_of_patch_sims(self) -> bool:
        return self._compute_mean_of_patch_sims

    def set_weight_patch_sims_in_combine(self, b: bool) -> None:
        self._weight_patch_sims_in_combine = bool(b)

    def weight_patch_sims_in_combine(self) -> bool:
        return self._weight_patch_sims_in_combine

    def set_use_mask_for_patch_weighting(self, b: bool) -> None:
        self._use_mask_for_weighting = bool(b)

    def use_mask_for_patch_weighting(self) -> bool:
        return self._use_mask_for_weighting

    def set_use_mask_for_patch_stats(self, b: bool) -> None:
        self._use_mask_for_patch_stats = bool(b)

    def use_mask_for_patch_stats(self) -> bool:
        return self._use_mask_for_patch_stats

    def set_normalize_weights_as_prob(self, b: bool) -> None:
        self._normalize_weights_as_prob = bool(b)

    def normalize_weights_as_prob(self) -> bool:
        return self._normalize_weights_as_prob

    def set_choose_rand_patches(self, b: bool) -> None:
        self._choose_rand_patches = bool(b)
        if not b and not self._do_not_update_patch_inds_to_use:
            self._push_subset(None)

    def choose_rand_patches(self) -> bool:
        return self._choose_rand_patches

    def set_num_rand_patches(self, n: int) -> None:
        self._num_rand_patches = int(n)

    def num_rand_patches(self) -> int:
        return self._num_rand_patches

    def set_rand_patch_min_pixels_sep(self, sep: float) -> None:
        self._rand_patch_min_pixels_sep = float(sep)

    def rand_patch_min_pixels_sep(self) -> float:
        return self._rand_patch_min_pixels_sep

    def seed_rand_patches(self, seed: int) -> None:
        """The reference seeds from std::random_device (SeedRNGEngWithRandDev); tests want reproducible draws."""
        self._rng = np.random.default_rng(seed)

    def set_patches_to_use(self, patch_inds) -> None:
        """set_patches_to_use (xregImgSimMetric2DPatchCommon.cpp:231-235): the metric is evaluated over this local
        list of global patch indices, in list order, until reset_patches_to_use()."""
        self._patch_inds_to_use = np.ascontiguousarray(patch_inds, dtype=np.uint64).reshape(-1)
        self._do_not_update_patch_inds_to_use = True
        self._push_subset(self._patch_inds_to_use)

    def reset_patches_to_use(self) -> None:
        self._patch_inds_to_use = None
        self._do_not_update_patch_inds_to_use = False
        self._push_subset(None)

    def patch_inds_to_use(self) -> Optional[np.ndarray]:
        return self._patch_inds_to_use

    def patch_indices_to_use(self) -> np.ndarray:
        """patch_indices_to_use (xregImgSimMetric2DPatchCommon.cpp:413-493): every patch, or num_rand_patches indices
        drawn (with replacement) from the discrete distribution of the patch weights, rejecting candidates closer
        than the minimum separation to an accepted one."""
        n_p = self.num_patches()
        if not self._choose_rand_patches:
            return np.arange(n_p, dtype=np.uint64)
        if not self._num_rand_patches < n_p:
            raise _lib.XregError("patch_indices_to_use: num_rand_patches must be smaller than the number of patches")
        w = self.compute_weights()
        p = None if w is None else (w.astype(np.float64) / float(np.sum(w, dtype=np.float64)))
        sep = self._rand_patch_min_pixels_sep
        min_sep = float(np.sqrt(2.0 * self._patch_radius * self._patch_radius)) if sep < 0 else sep
        check_sep = min_sep > 1.0e-8
        r, st = self._patch_radius, self._patch_stride
        ncc = (self._fixed.shape[1] - 1 - 2 * r) // st + 1
        inds, centres = [], []
        while len(inds) < self._num_rand_patches:
            k = int(self._rng.choice(n_p, p=p))
            if check_sep:
                c = np.array([r + (k // ncc) * st, r + (k % ncc) * st], dtype=np.float64)
                if any(np.linalg.norm(e - c) < min_sep for e in centres):
                    continue
                centres.append(c)
            inds.append(k)
        return np.asarray(inds, dtype=np.uint64)

    def _push_subset(self, inds: Optional[np.ndarray]) -> None:
        if inds is None or len(inds) == 0:
            check(self._lib.xrc_sm_set_patch_subset(self.handle, None, 0))
        else:
            a = np.ascontiguousarray(inds, dtype=np.uint64)
            check(self._lib.xrc_sm_set_patch_subset(self.handle, a.ctypes.data_as(C.POINTER(C.c_uint64)), a.size))

    def _pre_compute(self) -> None:
        # ImgSimMetric2DPatchNCCCPU::compute, xregImgSimMetric2DPatchNCCCPU.cpp:97-100: a fresh draw per compute()
        if self._choose_rand_patches and not self._do_not_update_patch_inds_to_use:
            self._patch_inds_to_use = self.patch_indices_to_use()
            self._push_subset(self._patch_inds_to_use)

    COMBINE_MODES = {"reference": 0, "reference-serial": 1, "f64": 2}

    def set_combine_mode(self, mode: str) -> None:
        """How the per-patch values become the image score (include/xreg_cuda.h, XRC_COMBINE_*): "reference"
        (default) reproduces the reference's sequential f32 sum and f32 total weight bit for bit
        (xregImgSimMetric2DPatchNCCCPU.cpp:262-287), "reference-serial" is the literal one-thread loop, "f64" sums in
        double (closest to exact arithmetic; differs from the CPU class by its f32 accumulation error)."""
        check(self._lib.xrc_sm_set_combine_mode(self.handle, self.COMBINE_MODES[mode]))

    def set_wgt_img(self, wgt_img: Optional[np.ndarray]) -> None:
        self._wgt_img = None if wgt_img is None else np.ascontiguousarray(wgt_img, dtype=f32)
        if self._allocated:
            self._push_patch_params()

    def num_patches(self) -> int:
        r, s = self._patch_radius, self._patch_stride
        rows, cols = self._fixed.shape
        return ((rows - 1 - 2 * r) // s + 1) * ((cols - 1 - 2 * r) // s + 1)

    def compute_weights(self) -> Optional[np.ndarray]:
        """ImgSimMetric2DPatchCommon::compute_weights (xregImgSimMetric2DPatchCommon.cpp:309-410).
        Returns None when every weight stays 1."""
        mask = self._mask
        use_mask_wgts = self._use_mask_for_weighting and mask is not None
        if self._wgt_img is None and not use_mask_wgts:
            return None
        r, s = self._patch_radius, self._patch_stride
        rows, cols = self._fixed.shape
        cr = np.arange(r, rows - r, s)
        cc = np.arange(r, cols - r, s)
        if self._wgt_img is not None:
            w = self._wgt_img[np.ix_(cr, cc)].astype(f32)
            if use_mask_wgts:
                w = np.where(mask[np.ix_(cr, cc)] != 0, w, f32(0)).astype(f32)
        else:
            d = 2 * r + 1
            ii = np.zeros((rows + 1, cols + 1), dtype=np.int64)
            ii[1:, 1:] = np.cumsum(np.cumsum((mask != 0).astype(np.int64), axis=0), axis=1)
            r0, c0 = cr - r, cc - r
            cnt = (ii[np.ix_(r0 + d, c0 + d)] - ii[np.ix_(r0, c0 + d)] - ii[np.ix_(r0 + d, c0)] + ii[np.ix_(r0, c0)])
            w = (cnt.astype(f32) / f32(d * d)).astype(f32)
        w = np.ascontiguousarray(w.reshape(-1), dtype=f32)
        if self._normalize_weights_as_prob:
            ws = np.cumsum(w, dtype=f32)[-1]  # sequential f32 sum, as the reference's loop
            w = (w / ws).astype(f32)
        return w

    def _push_patch_params(self) -> None:
        self._weights = self.compute_weights()
        wp, n = None, 0
        if self._weights is not None:
            wp, n = self._weights.ctypes.data_as(C.POINTER(C.c_float)), self._weights.size
        check(self._lib.xrc_sm_set_patch_params(self.handle, self._patch_radius, self._patch_stride,
                                                int(self._compute_mean_of_patch_sims),
                                                int(self._weight_patch_sims_in_combine),
                                                int(self._use_mask_for_patch_stats), wp, n))


class ImgSimMetric2DNCCCUDA(ImgSimMetric2D):
    """Replaces ImgSimMetric2DNCCOCL / mirrors ImgSimMetric2DNCCCPU (xregImgSimMetric2DNCCCPU.cpp).
    Unlike the CPU class the moving-image buffer is left untouched (the CPU class
    overwrites it with zero-mean images, xregImgSimMetric2DNCCCPU.h:36)."""

    KIND = _lib.SM_NCC


class ImgSimMetric2DSSDCUDA(ImgSimMetric2D):
    """Replaces ImgSimMetric2DSSDOCL / mirrors ImgSimMetric2DSSDCPU (xregImgSimMetric2DSSDCPU.cpp:62-110):
    sum((fixed - moving)^2) / num_pixels, both images taken as zero outside the mask.  The moving-image
    buffer is left untouched (the CPU class zeroes its masked pixels in place)."""

    KIND = _lib.SM_SSD


class ImgSimMetric2DGradNCCCUDA(ImgSimMetric2D, ImgSimMetric2DGradImgParamInterface):
    """ImgSimMetric2DGradNCCCPU (xregImgSimMetric2DGradNCCCPU.cpp:29-65)."""

    KIND = _lib.SM_GRAD_NCC


class ImgSimMetric2DPatchNCCCUDA(ImgSimMetric2D, ImgSimMetric2DPatchCommon):
    """ImgSimMetric2DPatchNCCCPU (xregImgSimMetric2DPatchNCCCPU.cpp)."""

    KIND = _lib.SM_PATCH_NCC
    _pre_compute = ImgSimMetric2DPatchCommon._pre_compute   # (the first base's hook is a no-op)

    def __init__(self, ctx: Context):
        ImgSimMetric2D.__init__(self, ctx)
        self._init_patch_common()

    def _pre_allocate(self) -> None:
        self._push_patch_params()

    def _mask_changed(self) -> None:
        if self._allocated:
            self._push_patch_params()


class ImgSimMetric2DPatchGradNCCCUDA(ImgSimMetric2D, ImgSimMetric2DPatchCommon, ImgSimMetric2DGradImgParamInterface):
    """ImgSimMetric2DPatchGradNCCCPU (xregImgSimMetric2DPatchGradNCCCPU.cpp:34-253)."""

    KIND = _lib.SM_PATCH_GRAD_NCC
    _pre_compute = ImgSimMetric2DPatchCommon._pre_compute

    def __init__(self, ctx: Context):
        ImgSimMetric2D.__init__(self, ctx)
        self._init_patch_common()

    def _pre_allocate(self) -> None:
        self._push_patch_params()

    def _mask_changed(self) -> None:
        if self._allocated:
            self._push_patch_params()


class ImgSimMetric2DCombineMean:
    """ImgSimMetric2DCombineMean (xregImgSimMetric2DCombine.cpp:67-86): host mean over views."""

    def __init__(self):
        self._sims: List[ImgSimMetric2D] = []
        self._sim_vals = np.zeros(0, dtype=f32)

    def set_sim_metrics(self, sims: Sequence[ImgSimMetric2D]) -> None:
        self._sims = list(sims)

    def compute(self) -> None:
        n = self._sims[0].num_moving_images()
        acc = np.zeros(n, dtype=f32)
        for s in self._sims:
            acc = (acc + s.sim_vals()[:n]).astype(f32)
        self._sim_vals = (acc / f32(len(self._sims))).astype(f32)

    def sim_vals(self) -> np.ndarray:
        return self._sim_vals

    def sim_val(self, i: int) -> float:
        return float(self._sim_vals[i])


def eval_batch(rc: RayCasterLineIntCUDA, sims: Sequence[ImgSimMetric2D], n_per_view: int, vol_idx: int = 0) -> np.ndarray:
    """One obj_fn evaluation (xregIntensity2D3DRegi.cpp:571-696): DRRs for all views,
    every view's metric, one gather.  Returns (n_views, n_per_view) float32."""
    lib = _lib.load()
    rc._flush()
    n_views = len(sims)
    arr = (C.c_void_p * n_views)(*[s.handle for s in sims])
    out = np.zeros((n_views, n_per_view), dtype=f32)
    check(lib.xrc_eval_batch(rc.handle, int(vol_idx), arr, n_views, int(n_per_view),
                             out.ctypes.data_as(C.POINTER(C.c_float))))
    for v, s in enumerate(sims):
        s._sim_vals[:n_per_view] = out[v]
    return out
