"""RayCasterLineIntCUDA: host-side mirror of xreg::RayCaster +
RayCastLineIntParamInterface (lib/ray_cast/xregRayCastInterface.h:43-434,575-591)
over the C ABI.  Same method names, argument meaning and error behaviour as the
reference interface; every method forwards to libxreg_cuda.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check
from .geometry import CameraModel, Volume, f32, to12


class Context:
    """One CUDA device + stream (replaces the OpenCL context/queue pair chosen by
    --ocl-id, lib/common/xregProgOptUtils.cpp:1712-1725)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = _lib.load()
        h = C.c_void_p()
        if stream is None:
            check(self._lib.xrc_ctx_create(int(device), C.byref(h)))
        else:
            check(self._lib.xrc_ctx_create_on_stream(int(device), C.c_void_p(stream), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def synchronize(self) -> None:
        check(self._lib.xrc_ctx_synchronize(self.handle))

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        check(self._lib.xrc_ctx_stream(self.handle, C.byref(s)))
        return int(s.value or 0)

    def close(self) -> None:
        if self.handle:
            self._lib.xrc_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RayCasterLineIntCUDA:
    # RayCaster::InterpMethod / ProjPixelStoreMethod / RayCastLineIntKernel
    kRAY_CAST_INTERP_LINEAR, kRAY_CAST_INTERP_NN, kRAY_CAST_INTERP_SINC, kRAY_CAST_INTERP_BSPLINE = 0, 1, 2, 3
    kRAY_CAST_PIXEL_REPLACE, kRAY_CAST_PIXEL_ACCUM = 0, 1
    kRAY_CAST_LINE_INT_SUM_KERNEL, kRAY_CAST_LINE_INT_MAX_KERNEL = 0, 1

    def __init__(self, ctx: Context, layout: str = "default"):
        self._lib = _lib.load()
        self.ctx = ctx
        h = C.c_void_p()
        check(self._lib.xrc_rc_create(ctx.handle, C.byref(h)))
        self.handle = h
        check(self._lib.xrc_rc_set_layout(self.handle, _lib.LAYOUT_NAMES[layout]))
        self._vols: List[Volume] = []
        self._cams: List[CameraModel] = []
        self._num_projs = 0
        self._max_num_projs = 0
        self._xforms: List[np.ndarray] = []          # xforms_cam_to_itk_phys_
        self._cam_model_for_proj: List[int] = []     # cam_model_for_proj_
        self._ray_step_size = 1.0                     # xregRayCastInterface.h:390
        self._interp_method = self.kRAY_CAST_INTERP_LINEAR
        self._proj_store_meth = self.kRAY_CAST_PIXEL_REPLACE
        self._kernel_id = self.kRAY_CAST_LINE_INT_SUM_KERNEL
        self._default_bg = 0.0
        self._use_bg_projs = False
        self._bg_projs: List[np.ndarray] = []
        self._poses_dirty = True
        self._params_dirty = True
        self._resources_allocated = False

    # ---- volumes / cameras -------------------------------------------------
    def set_volume(self, vol: Volume) -> None:
        self.set_volumes([vol])

    def set_volumes(self, vols: Sequence[Volume]) -> None:
        self._vols = list(vols)
        self.vols_changed()

    def set_volumes_hu(self, vols: Sequence[Volume], hu_lower: float = -1000.0) -> None:
        """Volumes in Hounsfield units: converted to linear attenuation on the device while loading
        (HUToLinAtt, lib/image/xregHUToLinAtt.cpp:45-69), bit-identical to converting on the host first."""
        self._vols = list(vols)
        n = len(self._vols)
        ptrs = (C.POINTER(C.c_float) * n)(*[v.data.ctypes.data_as(C.POINTER(C.c_float)) for v in self._vols])
        dims = ((C.c_uint64 * 3) * n)(*[(C.c_uint64 * 3)(*v.dims) for v in self._vols])
        xf = ((C.c_float * 12) * n)(*[(C.c_float * 12)(*[float(t) for t in v.idx_to_phys()]) for v in self._vols])
        check(self._lib.xrc_rc_set_volumes_hu(self.handle, n, ptrs, dims, xf, float(hu_lower)))

    def num_vols(self) -> int:
        return len(self._vols)

    def vols_changed(self) -> None:
        n = len(self._vols)
        ptrs = (C.POINTER(C.c_float) * n)(*[v.data.ctypes.data_as(C.POINTER(C.c_float)) for v in self._vols])
        dims = ((C.c_uint64 * 3) * n)(*[(C.c_uint64 * 3)(*v.dims) for v in self._vols])
        xf = ((C.c_float * 12) * n)(*[(C.c_float * 12)(*[float(t) for t in v.idx_to_phys()]) for v in self._vols])
        check(self._lib.xrc_rc_set_volumes(self.handle, n, ptrs, dims, xf))

    def set_camera_model(self, cam: CameraModel) -> None:
        self.set_camera_models([cam])

    def set_camera_models(self, cams: Sequence[CameraModel]) -> None:
        self._cams = list(cams)
        self.camera_models_changed()

    def camera_models_changed(self) -> None:
        arr = (_lib.XrcCam * len(self._cams))(*[c.to_xrc() for c in self._cams])
        check(self._lib.xrc_rc_set_cameras(self.handle, len(self._cams), arr))

    def num_camera_models(self) -> int:
        return len(self._cams)

    def camera_models(self) -> List[CameraModel]:
        return self._cams

    def camera_model(self, cam_idx: int = 0) -> CameraModel:
        return self._cams[cam_idx]

    # ---- projections / poses -----------------------------------------------
    def set_num_projs(self, num_projs: int) -> None:
        """xregRayCastInterface.cpp:131-139; capacity is fixed at allocate_resources()."""
        if self._resources_allocated:
            check(self._lib.xrc_rc_set_num_projs(self.handle, int(num_projs)))
        self._num_projs = int(num_projs)
        self._cam_model_for_proj = (self._cam_model_for_proj + [0] * num_projs)[:num_projs]
        ident = np.eye(4, dtype=f32)
        self._xforms = (self._xforms + [ident.copy() for _ in range(num_projs)])[:num_projs]
        self._poses_dirty = True

    def num_projs(self) -> int:
        return self._num_projs

    def max_num_projs(self) -> int:
        return self._max_num_projs

    def set_proj_cam_model(self, proj_idx: int, cam_idx: int) -> None:
        self._cam_model_for_proj[proj_idx] = int(cam_idx)
        self._poses_dirty = True

    def camera_model_proj_associations(self) -> List[int]:
        return self._cam_model_for_proj

    def set_camera_model_proj_associations(self, assoc: Sequence[int]) -> None:
        if len(assoc) != self._num_projs:
            raise _lib.XregError("set_camera_model_proj_associations: size must equal num_projs")
        self._cam_model_for_proj = [int(a) for a in assoc]
        self._poses_dirty = True

    def set_xforms_cam_to_itk_phys(self, xforms: Sequence[np.ndarray]) -> None:
        if len(xforms) != self._num_projs:
            raise _lib.XregError("set_xforms_cam_to_itk_phys: size must equal num_projs")
        self._xforms = [np.asarray(x, dtype=f32).reshape(4, 4).copy() for x in xforms]
        self._poses_dirty = True

    def xforms_cam_to_itk_phys(self) -> List[np.ndarray]:
        return self._xforms

    def xform_cam_to_itk_phys(self, proj_idx: int) -> np.ndarray:
        """Mutable reference like the C++ accessor; marks the pose list dirty."""
        self._poses_dirty = True
        return self._xforms[proj_idx]

    def distribute_xforms_among_cam_models(self, xforms: Sequence[np.ndarray]) -> None:
        """xregRayCastInterface.cpp:97-114: camera-major replication."""
        n_passed, n_cams = len(xforms), self.num_camera_models()
        if n_passed * n_cams != self._num_projs:
            raise _lib.XregError("distribute_xforms_among_cam_models: n_xforms * n_cams must equal num_projs")
        g = 0
        for cam_idx in range(n_cams):
            for p in range(n_passed):
                self._xforms[g] = np.asarray(xforms[p], dtype=f32).reshape(4, 4).copy()
                self._cam_model_for_proj[g] = cam_idx
                g += 1
        self._poses_dirty = True

    def distribute_xform_among_cam_models(self, xform: np.ndarray) -> None:
        self.distribute_xforms_among_cam_models([xform])

    def post_multiply_all_xforms(self, post_xform: np.ndarray) -> None:
        self._xforms = [(x @ np.asarray(post_xform, dtype=f32)).astype(f32) for x in self._xforms]
        self._poses_dirty = True

    def pre_multiply_all_xforms(self, pre_xform: np.ndarray) -> None:
        self._xforms = [(np.asarray(pre_xform, dtype=f32) @ x).astype(f32) for x in self._xforms]
        self._poses_dirty = True

    # ---- parameters --------------------------------------------------------
    def set_ray_step_size(self, step_size: float) -> None:
        self._ray_step_size = float(step_size)
        self._params_dirty = True

    def ray_step_size(self) -> float:
        return self._ray_step_size

    def set_interp_method(self, m: int) -> None:
        self._interp_method = int(m)
        self._params_dirty = True

    def interp_method(self) -> int:
        return self._interp_method

    def use_linear_interp(self) -> None:
        self.set_interp_method(self.kRAY_CAST_INTERP_LINEAR)

    def use_nn_interp(self) -> None:
        self.set_interp_method(self.kRAY_CAST_INTERP_NN)

    def use_sinc_interp(self) -> None:
        self.set_interp_method(self.kRAY_CAST_INTERP_SINC)

    def use_bspline_interp(self) -> None:
        self.set_interp_method(self.kRAY_CAST_INTERP_BSPLINE)

    def set_proj_store_method(self, m: int) -> None:
        self._proj_store_meth = int(m)
        self._params_dirty = True

    def proj_store_method(self) -> int:
        return self._proj_store_meth

    def use_proj_store_replace_method(self) -> None:
        self.set_proj_store_method(self.kRAY_CAST_PIXEL_REPLACE)

    def use_proj_store_accum_method(self) -> None:
        self.set_proj_store_method(self.kRAY_CAST_PIXEL_ACCUM)

    def kernel_id(self) -> int:
        return self._kernel_id

    def set_kernel_id(self, k: int) -> None:
        self._kernel_id = int(k)
        self._params_dirty = True

    def default_bg_pixel_val(self) -> float:
        return self._default_bg

    def set_default_bg_pixel_val(self, v: float) -> None:
        self._default_bg = float(v)
        self._params_dirty = True

    def set_use_bg_projs(self, use: bool) -> None:
        self._use_bg_projs = bool(use)
        self._push_bg()

    def use_bg_projs(self) -> bool:
        return self._use_bg_projs

    def set_bg_proj(self, proj: np.ndarray, use_bg_projs: bool = True) -> None:
        self.set_bg_projs([proj], use_bg_projs)

    def set_bg_projs(self, projs: Sequence[np.ndarray], use_bg_projs: bool = True) -> None:
        self._bg_projs = [np.ascontiguousarray(p, dtype=f32) for p in projs]
        self._use_bg_projs = bool(use_bg_projs)
        self._push_bg(upload=True)

    def _push_bg(self, upload: bool = False) -> None:
        if self._use_bg_projs:
            if len(self._bg_projs) != self.num_camera_models():
                raise _lib.XregError("background projections: need one per camera model (xregRayCastBaseCPU.cpp:135)")
            if upload:
                arr = (C.POINTER(C.c_float) * len(self._bg_projs))(
                    *[b.ctypes.data_as(C.POINTER(C.c_float)) for b in self._bg_projs])
                check(self._lib.xrc_rc_set_bg_projs(self.handle, arr, 1))
            else:
                check(self._lib.xrc_rc_set_bg_projs(self.handle, None, 1))
        else:
            check(self._lib.xrc_rc_set_bg_projs(self.handle, None, 0))

    def set_layout_order(self, order: int) -> None:
        check(self._lib.xrc_rc_set_cta_order(self.handle, int(order)))

    # ---- resources / compute -----------------------------------------------
    def allocate_resources(self) -> None:
        """RayCaster::allocate_resources (xregRayCastInterface.cpp:262-270): capacity = num_projs now."""
        if self._num_projs <= 0:
            raise _lib.XregError("allocate_resources: set_num_projs first")
        check(self._lib.xrc_rc_allocate(self.handle, self._num_projs))
        self._max_num_projs = self._num_projs
        self._resources_allocated = True
        self._poses_dirty = True
        self._params_dirty = True

    def max_num_projs_possible(self) -> int:
        n = C.c_uint64()
        check(self._lib.xrc_rc_max_projs_possible(self.handle, C.byref(n)))
        return int(n.value)

    def use_other_proj_buf(self, other: "RayCasterLineIntCUDA") -> None:
        check(self._lib.xrc_rc_use_other_proj_buf(self.handle, other.handle))

    def _flush_params(self) -> None:
        if self._params_dirty:
            check(self._lib.xrc_rc_set_params(self.handle, self._ray_step_size, self._interp_method, self._kernel_id,
                                              self._proj_store_meth, self._default_bg))
            self._params_dirty = False

    def _flush(self) -> None:
        self._flush_params()
        if self._poses_dirty and self._num_projs:
            poses = to12(np.stack(self._xforms[: self._num_projs]))
            idx = np.asarray(self._cam_model_for_proj[: self._num_projs], dtype=np.uint32)
            check(self._lib.xrc_rc_set_poses(self.handle, self._num_projs, poses.ctypes.data_as(C.POINTER(C.c_float)),
                                             idx.ctypes.data_as(C.POINTER(C.c_uint32))))
            self._poses_dirty = False

    def set_poses_array(self, poses12: np.ndarray, cam_idx: Optional[np.ndarray] = None) -> None:
        """Fast path used by the registration loop: (n, 12) float32 array straight to the ABI."""
        poses12 = np.ascontiguousarray(poses12, dtype=f32).reshape(-1, 12)
        n = poses12.shape[0]
        if n != self._num_projs:
            raise _lib.XregError("set_poses_array: pose count must equal num_projs")
        ip = None
        if cam_idx is not None:
            cam_idx = np.ascontiguousarray(cam_idx, dtype=np.uint32)
            ip = cam_idx.ctypes.data_as(C.POINTER(C.c_uint32))
        check(self._lib.xrc_rc_set_poses(self.handle, n, poses12.ctypes.data_as(C.POINTER(C.c_float)), ip))
        self._xforms = [np.vstack([p.reshape(3, 4), np.array([[0, 0, 0, 1]], dtype=f32)]) for p in poses12]
        self._cam_model_for_proj = [0] * n if cam_idx is None else [int(c) for c in cam_idx]
        self._poses_dirty = False

    def set_poses_device(self, dev_poses_ptr: int, n: int, dev_cam_idx_ptr: int = 0,
                         host_mirror: Optional[np.ndarray] = None, host_cam_idx: Optional[np.ndarray] = None) -> None:
        """Poses already resident on the device (n x 12 float32, optional n x uint32 camera ids).  host_mirror: the same
        values on the host, so that only the principal-axis stacks these poses need are built (without it: all three)."""
        if host_mirror is None:
            check(self._lib.xrc_rc_set_poses_device(self.handle, int(n), C.c_void_p(dev_poses_ptr),
                                                    C.c_void_p(dev_cam_idx_ptr) if dev_cam_idx_ptr else None))
        else:
            hm = np.ascontiguousarray(host_mirror, dtype=f32).reshape(-1, 12)
            hc = None
            if dev_cam_idx_ptr:
                hc_arr = np.ascontiguousarray(host_cam_idx, dtype=np.uint32)
                hc = hc_arr.ctypes.data_as(C.POINTER(C.c_uint32))
            check(self._lib.xrc_rc_set_poses_device_mirrored(
                self.handle, int(n), C.c_void_p(dev_poses_ptr), C.c_void_p(dev_cam_idx_ptr) if dev_cam_idx_ptr else None,
                hm.ctypes.data_as(C.POINTER(C.c_float)), hc))
        self._poses_dirty = False

    # ---- multi-GPU tile sharding (one process per GPU; include/xreg_cuda.h "Tile-sharded objective") ----
    def peer_export(self) -> bytes:
        """CUDA IPC handle of this ray caster's projection buffer (after allocate_resources)."""
        h = (C.c_uint8 * 64)()
        check(self._lib.xrc_rc_peer_export(self.handle, h))
        return bytes(h)

    def peer_attach(self, n_ranks: int, rank: int, handles: Sequence[bytes]) -> None:
        """Open the projection buffers of all ranks (their peer_export() handles, in rank order)."""
        blob = (C.c_uint8 * (64 * n_ranks)).from_buffer_copy(b"".join(bytes(h) for h in handles))
        check(self._lib.xrc_rc_peer_attach(self.handle, int(n_ranks), int(rank), blob))

    def peer_detach(self) -> None:
        check(self._lib.xrc_rc_peer_detach(self.handle))

    def compute_tiles(self, vol_idx: int = 0) -> None:
        """compute() for this rank's detector tiles of ALL current projections, each written into its owner's buffer."""
        if not self._resources_allocated:
            raise _lib.XregError("compute_tiles: resources not allocated")
        self._flush()
        check(self._lib.xrc_rc_compute_tiles(self.handle, int(vol_idx)))

    def plan_tiles(self, vol_idx: int = 0, n_ranks: int = 0) -> List[int]:
        """(Re)balance the ranks' tile ranges on the current projections; returns the n_ranks + 1 bounds."""
        self._flush()
        check(self._lib.xrc_rc_plan_tiles(self.handle, int(vol_idx)))
        out = (C.c_uint32 * 9)()
        check(self._lib.xrc_rc_tile_plan(self.handle, out))
        return [int(v) for v in out[: (n_ranks + 1) if n_ranks else 9]]

    def plan_tiles_timed(self, rank_ms: Sequence[float], vol_idx: int = 0) -> List[int]:
        """Re-cut the ranks' tile ranges from the kernel times they measured under the current plan (the same list
        on every rank); returns the new bounds."""
        self._flush()
        ms = (C.c_float * len(rank_ms))(*[float(v) for v in rank_ms])
        check(self._lib.xrc_rc_plan_tiles_timed(self.handle, int(vol_idx), ms))
        out = (C.c_uint32 * 9)()
        check(self._lib.xrc_rc_tile_plan(self.handle, out))
        return [int(v) for v in out[: len(rank_ms) + 1]]

    def tile_samples(self, vol_idx: int = 0):
        """(algorithmic, fetched) trilinear samples of this rank's tiles for the current poses."""
        self._flush()
        a, f = C.c_uint64(), C.c_uint64()
        check(self._lib.xrc_rc_tile_samples(self.handle, int(vol_idx), C.byref(a), C.byref(f)))
        return int(a.value), int(f.value)

    def compute(self, vol_idx: int = 0) -> None:
        if not self._resources_allocated:
            raise _lib.XregError("compute: resources not allocated (xregRayCastLineIntCPU.cpp:296)")
        self._flush()
        check(self._lib.xrc_rc_compute(self.handle, int(vol_idx)))

    def proj(self, proj_idx: int) -> np.ndarray:
        """Copy of projection proj_idx (the reference returns a view of the host buffer)."""
        cam = self._cams[self._cam_model_for_proj[proj_idx] if proj_idx < len(self._cam_model_for_proj) else 0]
        out = np.empty((cam.num_det_rows, cam.num_det_cols), dtype=f32)
        check(self._lib.xrc_rc_read_projs(self.handle, int(proj_idx), 1, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    proj_ocv = proj

    def raw_host_pixel_buf(self) -> np.ndarray:
        """All current projections, image-major then row-major (xregRayCastBaseCPU.h:149-154)."""
        cam = self._cams[0]
        out = np.empty((self._num_projs, cam.num_det_rows, cam.num_det_cols), dtype=f32)
        if self._num_projs:
            check(self._lib.xrc_rc_read_projs(self.handle, 0, self._num_projs, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def use_external_host_pixel_buf(self, buf) -> None:
        raise _lib.UnsupportedOperationException(
            "use_external_host_pixel_buf: projections live in device memory; use raw_host_pixel_buf()/proj()")

    def device_buf(self) -> int:
        p = C.c_void_p()
        check(self._lib.xrc_rc_device_buf(self.handle, C.byref(p)))
        return int(p.value)

    def ray_info(self, vol_idx: int = 0, counts_only: bool = False):
        """(clip mask, samples per ray, total samples S) for the current poses
        (mask / per-ray arrays are None with counts_only)."""
        self._flush()
        S = C.c_uint64()
        if counts_only:
            check(self._lib.xrc_rc_ray_info(self.handle, int(vol_idx), None, None, C.byref(S)))
            return None, None, int(S.value)
        cam = self._cams[0]
        shape = (self._num_projs, cam.num_det_rows, cam.num_det_cols)
        mask = np.zeros(shape, np.uint8)
        steps = np.zeros(shape, np.uint32)
        check(self._lib.xrc_rc_ray_info(self.handle, int(vol_idx), mask.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        steps.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(S)))
        return mask, steps, int(S.value)

    def volume_bytes(self) -> int:
        """Device bytes of the volume representation right now (PAX stacks are built on demand)."""
        n = C.c_uint64()
        check(self._lib.xrc_rc_volume_bytes(self.handle, C.byref(n)))
        return int(n.value)

    def set_skip_empty(self, enable) -> None:
        """Empty-space trimming of the sum kernel (exact, default on); off only for measurement.  3: on, plus skipping of
        empty runs INSIDE a ray's range (structures far apart along the view direction; same bits)."""
        check(self._lib.xrc_rc_set_skip_empty(self.handle, int(enable) if not isinstance(enable, bool) else (1 if enable else 0)))

    def fetched_samples(self, vol_idx: int = 0) -> int:
        """Trilinear samples compute() actually fetches for the current poses (<= S of ray_info)."""
        self._flush()
        n = C.c_uint64()
        check(self._lib.xrc_rc_fetched_samples(self.handle, int(vol_idx), C.byref(n)))
        return int(n.value)

    def close(self) -> None:
        if self.handle:
            self._lib.xrc_rc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


kRAY_CAST_MAX_DEPTH = 1.0e37   # xregRayCastInterface.h:601


class RayCasterDepthCUDA(RayCasterLineIntCUDA):
    """RayCasterDepthCPU (lib/ray_cast/xregRayCastDepthCPU.{h,cpp}) + RayCasterCollisionParamInterface
    (xregRayCastInterface.h:436-475, .cpp:352-371): every pixel gets the depth -- distance from the pinhole in the camera
    frame -- of the first sample along its ray whose interpolated value reaches render_thresh(), refined by
    num_backtracking_steps() halvings of the step; combined with min on top of the background, which defaults to
    kRAY_CAST_MAX_DEPTH (the class's constructor, xregRayCastDepthCPU.cpp:231-234).  Linear or nearest-neighbour
    interpolation.  Same volumes / cameras / poses / store methods as the line-integral ray caster."""

    def __init__(self, ctx: Context, layout: str = "default"):
        super().__init__(ctx, layout)
        self.set_default_bg_pixel_val(kRAY_CAST_MAX_DEPTH)
        self._render_thresh = 150.0          # xregRayCastInterface.cpp:427-428
        self._num_backtracking_steps = 0

    def set_render_thresh(self, t: float) -> None:
        self._render_thresh = float(t)

    def render_thresh(self) -> float:
        return self._render_thresh

    def set_num_backtracking_steps(self, n: int) -> None:
        self._num_backtracking_steps = int(n)

    def num_backtracking_steps(self) -> int:
        return self._num_backtracking_steps

    def compute(self, vol_idx: int = 0) -> None:
        if not self._resources_allocated:
            raise _lib.XregError("compute: resources not allocated (xregRayCastDepthCPU.cpp:238)")
        self._flush()
        check(self._lib.xrc_rc_compute_depth(self.handle, int(vol_idx), self._render_thresh, self._num_backtracking_steps))
