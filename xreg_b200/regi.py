"""The batch objective of intensity-based 2D/3D registration and its multi-GPU sharding.

Intensity2D3DObjFn is what Intensity2D3DRegi::obj_fn does per optimiser iteration
(lib/regi/interfaces_2d_3d/xregIntensity2D3DRegi.cpp:571-696) for one moving volume:
distribute the population over the views (camera-major), ray cast, evaluate every
view's metric, average over views.  ShardedObjFn splits the population over the
ranks of a torch.distributed job (one process per GPU, volume replicated, no
data-path collective) and gathers only the per-pose scalars.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .geometry import CameraModel, Volume, f32, to12
from .ray_caster import Context, RayCasterLineIntCUDA
from .sim_metrics import (ImgSimMetric2D, ImgSimMetric2DGradNCCCUDA, ImgSimMetric2DNCCCUDA,
                          ImgSimMetric2DPatchGradNCCCUDA, ImgSimMetric2DPatchNCCCUDA, ImgSimMetric2DSSDCUDA, eval_batch)

METRICS = {
    "ssd": ImgSimMetric2DSSDCUDA,
    "ncc": ImgSimMetric2DNCCCUDA,
    "grad-ncc": ImgSimMetric2DGradNCCCUDA,
    "patch-ncc": ImgSimMetric2DPatchNCCCUDA,
    "patch-grad-ncc": ImgSimMetric2DPatchGradNCCCUDA,
}


def shard_bounds(n: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [begin, end) pose ranges per rank; the first n % world ranks
    take one extra pose (100 poses on 8 ranks -> 13,13,13,13,12,12,12,12)."""
    base, extra = divmod(int(n), int(world_size))
    out, b = [], 0
    for r in range(world_size):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


class Intensity2D3DObjFn:
    """Single-GPU objective: poses (n, 4, 4) or (n, 12) cam->volume-physical -> (n,) similarity."""

    def __init__(self, ctx: Context, vol: Volume, cams: Sequence[CameraModel], fixed_imgs: Sequence[np.ndarray],
                 metric: str = "patch-grad-ncc", max_pop: int = 100, patch_radius: Optional[int] = None,
                 patch_stride: int = 1, gauss_width: int = 5, masks: Optional[Sequence[Optional[np.ndarray]]] = None,
                 step_size: float = 1.0, layout: str = "default"):
        if len(cams) != len(fixed_imgs):
            raise _lib.XregError("need one fixed image per camera model / view")
        self.ctx = ctx
        self.n_views = len(cams)
        self.max_pop = int(max_pop)
        self.rc = RayCasterLineIntCUDA(ctx, layout=layout)
        # one moving volume, or several (multi-object registration: eval_objects)
        self.rc.set_volumes(list(vol) if isinstance(vol, (list, tuple)) else [vol])
        self.rc.set_camera_models(list(cams))
        self.rc.set_ray_step_size(step_size)
        # xregIntensity2D3DRegi.cpp:63-94: view-major buffer, metric v reads [v*pop, (v+1)*pop)
        self.rc.set_num_projs(self.max_pop * self.n_views)
        self.rc.allocate_resources()
        self.sims: List[ImgSimMetric2D] = []
        for v in range(self.n_views):
            sm = METRICS[metric](ctx)
            sm.set_num_moving_images(self.max_pop)
            sm.set_fixed_image(fixed_imgs[v])
            sm.set_mov_imgs_buf_from_ray_caster(self.rc, self.max_pop * v)
            if masks is not None and masks[v] is not None:
                sm.set_mask(masks[v])
            if hasattr(sm, "set_patch_radius") and patch_radius is not None:
                sm.set_patch_radius(patch_radius)
                sm.set_patch_stride(patch_stride)
            if hasattr(sm, "set_smooth_img_before_sobel_kernel_radius"):
                sm.set_smooth_img_before_sobel_kernel_radius(gauss_width)
            sm.allocate_resources()
            self.sims.append(sm)
        self._cur_pop = self.max_pop
        self._lib = _lib.load()
        self._sm_arr = (C.c_void_p * self.n_views)(*[sm.handle for sm in self.sims])

    def _set_pop(self, n: int) -> None:
        if n > self.max_pop:
            raise _lib.XregError("population larger than the allocated capacity")
        if n != self._cur_pop:
            self.rc.set_num_projs(n * self.n_views)
            for v, sm in enumerate(self.sims):
                sm.set_num_moving_images(n)
                sm.set_mov_imgs_buf_from_ray_caster(self.rc, n * v)
            self._cur_pop = n

    def __call__(self, poses: np.ndarray) -> np.ndarray:
        """One objective evaluation = one call into the library (xrc_obj_fn): distribute the poses over
        the views camera-major, ray cast, every view's metric, one gather, mean over views."""
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        if n == 0:
            return np.zeros(0, dtype=f32)
        self._set_pop(n)
        self.rc._flush_params()
        for sm in self.sims:
            sm._pre_compute()   # random patch subsets are drawn per evaluation
        out = np.empty(n, dtype=f32)
        per_view = np.empty((self.n_views, n), dtype=f32)
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn(self.rc.handle, 0, self._sm_arr, self.n_views, n, p12.ctypes.data_as(FP),
                                        out.ctypes.data_as(FP), per_view.ctypes.data_as(FP)))
        self.rc._poses_dirty = False  # the library now holds the distributed poses
        for v, sm in enumerate(self.sims):
            sm._sim_vals[:n] = per_view[v]
        return out

    def eval_units(self, poses: np.ndarray, first_unit: int, n_units: int) -> np.ndarray:
        """This device's part of a sharded objective (xrc_obj_fn_units): the per-view similarity values of the units
        [first_unit, first_unit + n_units) of the camera-major list u = view * n + pose, for all n poses given."""
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        out = np.empty(int(n_units), dtype=f32)
        if n_units == 0:
            return out
        if n > self.max_pop:
            raise _lib.XregError("population larger than the allocated capacity")
        self.rc._flush_params()
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn_units(self.rc.handle, 0, self._sm_arr, self.n_views, n, p12.ctypes.data_as(FP),
                                              int(first_unit), int(n_units), out.ctypes.data_as(FP)))
        self.rc._poses_dirty = False
        self._cur_pop = -1   # the library re-sized / re-bound the objects
        return out

    def enqueue_units(self, poses: np.ndarray, first_unit: int, n_units: int) -> None:
        """eval_units without the synchronise / read-back (xrc_obj_fn_units_enqueue): view v's values are left in the
        first entries of self.sims[v]'s device result vector, for a gather on the device (ShardedDeviceObjFn)."""
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        if n_units == 0:
            return
        self.rc._flush_params()   # the library checks the chunk against the allocated capacity
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn_units_enqueue(self.rc.handle, 0, self._sm_arr, self.n_views, n,
                                                      p12.ctypes.data_as(FP), int(first_unit), int(n_units)))
        self.rc._poses_dirty = False
        self._cur_pop = -1   # the library re-sized / re-bound the objects

    def enqueue_tiles_drr(self, poses: np.ndarray) -> int:
        """First half of the tile-sharded objective (xrc_obj_fn_tiles_enqueue_drr): all poses, this rank's tiles."""
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        self.rc._flush_params()
        _lib.check(self._lib.xrc_obj_fn_tiles_enqueue_drr(self.rc.handle, 0, self.n_views, n, p12.ctypes.data_as(C.POINTER(C.c_float))))
        self.rc._poses_dirty = False
        self._cur_pop = -1
        return n

    def enqueue_units_metrics(self, n: int, first_unit: int, n_units: int) -> None:
        """Second half (after the ranks' barrier): the metrics of the units this rank owns (xrc_obj_fn_units_enqueue_metrics)."""
        if n_units == 0:
            return
        for sm in self.sims:
            sm._pre_compute()
        _lib.check(self._lib.xrc_obj_fn_units_enqueue_metrics(self.rc.handle, self._sm_arr, self.n_views, int(n), int(first_unit),
                                                              int(n_units)))
        self._cur_pop = -1

    def eval_tiles(self, poses: np.ndarray) -> np.ndarray:
        """The whole tile-sharded evaluation in one library call (xrc_obj_fn_tiles; every rank calls it with the same
        poses after peer_attach): ray cast this rank's tiles into their owners' buffers, barrier, metrics of the units
        this rank owns, all-gather -- barrier and gather by the library's own kernel over the peer mappings, no NCCL."""
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        if n == 0:
            return np.zeros(0, dtype=f32)
        self.rc._flush_params()
        for sm in self.sims:
            sm._pre_compute()
        out = np.empty(n, dtype=f32)
        per_view = np.empty((self.n_views, n), dtype=f32)
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn_tiles(self.rc.handle, 0, self._sm_arr, self.n_views, n, p12.ctypes.data_as(FP),
                                              out.ctypes.data_as(FP), per_view.ctypes.data_as(FP)))
        self.rc._poses_dirty = False
        self._cur_pop = -1   # the library re-sized / re-bound the objects
        self.per_view = per_view
        return out

    def close(self) -> None:
        """Destroy the metrics and the ray caster (before their Context is closed)."""
        for sm in self.sims:
            sm.close()
        self.sims = []
        self.rc.close()

    def eval_objects(self, poses_per_object: Sequence[np.ndarray], vol_inds: Optional[Sequence[int]] = None,
                     use_bg_projs: bool = False) -> np.ndarray:
        """Multi-object objective (Intensity2D3DRegi::obj_fn's loop over volumes, xregIntensity2D3DRegi.cpp:594-629):
        object j's population poses_per_object[j] (n, 4, 4) is ray cast through volume vol_inds[j] (default j) into the
        same projections -- first object REPLACE (on the background projections if use_bg_projs), the others ACCUM --
        then the metrics.  One library call (xrc_obj_fn_objects)."""
        n_objs = len(poses_per_object)
        p12 = np.ascontiguousarray(np.stack([to12(p) if np.asarray(p).ndim == 3 else np.asarray(p, f32).reshape(-1, 12)
                                             for p in poses_per_object]), dtype=f32)
        n = p12.shape[1]
        if n == 0:
            return np.zeros(0, dtype=f32)
        self._set_pop(n)
        self.rc._flush_params()
        vi = np.ascontiguousarray(np.arange(n_objs) if vol_inds is None else vol_inds, dtype=np.uint32)
        out = np.empty(n, dtype=f32)
        per_view = np.empty((self.n_views, n), dtype=f32)
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn_objects(self.rc.handle, n_objs, vi.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                self._sm_arr, self.n_views, n, p12.ctypes.data_as(FP),
                                                1 if use_bg_projs else 0, out.ctypes.data_as(FP),
                                                per_view.ctypes.data_as(FP)))
        self.rc._poses_dirty = False
        for v, sm in enumerate(self.sims):
            sm._sim_vals[:n] = per_view[v]
        return out

    def eval_se3(self, params: np.ndarray, pre: Optional[np.ndarray] = None, post: Optional[np.ndarray] = None,
                 penalty: Optional["_lib.XrcSe3Penalty"] = None) -> np.ndarray:
        """Objective from optimiser variables (SE3OptVarsLieAlg, xregSE3OptVars.cpp:128-137):
        pose_p = pre * ExpSE3(params_p) * post (xregIntensity2D3DRegi.cpp:1049-1071), composed inside the library;
        with `penalty` (se3_penalty()) the regulariser of xregIntensity2D3DRegi.cpp:653-688 is added in the same call
        (self.last_penalty holds the unscaled values)."""
        x = np.ascontiguousarray(params, dtype=f32).reshape(-1, 6)
        n = x.shape[0]
        if n == 0:
            return np.zeros(0, dtype=f32)
        self._set_pop(n)
        self.rc._flush_params()
        for sm in self.sims:
            sm._pre_compute()
        out = np.empty(n, dtype=f32)
        FP = C.POINTER(C.c_float)
        pre_a = None if pre is None else to12(np.asarray(pre, f32)[None])
        post_a = None if post is None else to12(np.asarray(post, f32)[None])
        pre12 = None if pre_a is None else pre_a.ctypes.data_as(FP)
        post12 = None if post_a is None else post_a.ctypes.data_as(FP)
        if penalty is None:
            _lib.check(self._lib.xrc_obj_fn_se3(self.rc.handle, 0, self._sm_arr, self.n_views, n, x.ctypes.data_as(FP),
                                                pre12, post12, out.ctypes.data_as(FP), None))
        else:
            self.last_penalty = np.empty(n, dtype=f32)
            _lib.check(self._lib.xrc_obj_fn_se3_pen(self.rc.handle, 0, self._sm_arr, self.n_views, n, x.ctypes.data_as(FP),
                                                    pre12, post12, C.byref(penalty), out.ctypes.data_as(FP), None,
                                                    self.last_penalty.ctypes.data_as(FP)))
        self.rc._poses_dirty = False
        return out


def se3_penalty(rot_mean: float, rot_std: float, trans_mean: float, trans_std: float,
                inter_frame: Optional[np.ndarray] = None, init_cam_to_vol: Optional[np.ndarray] = None,
                inter_wrt_vol: bool = True, img_sim_coeff: Optional[float] = None,
                penalty_coeff: Optional[float] = None) -> "_lib.XrcSe3Penalty":
    """Regi2D3DPenaltyFnSE3Mag with FoldNormDist(rot_mean, rot_std) / FoldNormDist(trans_mean, trans_std) for one object
    (xregRegi2D3DPenaltyFnSE3Mag.cpp, xregFoldNormDist.cpp); coefficients as set_img_sim_penalty_coefs."""
    p = _lib.XrcSe3Penalty()
    p.rot_mean, p.rot_std, p.trans_mean, p.trans_std = float(rot_mean), float(rot_std), float(trans_mean), float(trans_std)
    p.use_coeffs = 0 if img_sim_coeff is None and penalty_coeff is None else 1
    p.img_sim_coeff = 1.0 if img_sim_coeff is None else float(img_sim_coeff)
    p.penalty_coeff = 1.0 if penalty_coeff is None else float(penalty_coeff)
    p.inter_wrt_vol = 1 if inter_wrt_vol else 0
    eye = np.eye(4, dtype=f32)
    p.inter_frame[:] = [float(v) for v in to12((eye if inter_frame is None else np.asarray(inter_frame, f32))[None])[0]]
    p.init_cam_to_vol[:] = [float(v) for v in to12((eye if init_cam_to_vol is None else np.asarray(init_cam_to_vol, f32))[None])[0]]
    return p


def se3_mag_penalty(pen: "_lib.XrcSe3Penalty", poses: np.ndarray) -> np.ndarray:
    """reg_vals of Regi2D3DPenaltyFnSE3Mag::compute for the poses (n, 4, 4) / (n, 12) handed to the ray caster (host only)."""
    p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
    out = np.empty(p12.shape[0], dtype=f32)
    FP = C.POINTER(C.c_float)
    _lib.check(_lib.load().xrc_se3_mag_penalty(C.byref(pen), p12.shape[0], p12.ctypes.data_as(FP), out.ctypes.data_as(FP)))
    return out


def multi_device_share(n_dev: int, n_views: int, n_poses: int, dev: int, view: int) -> Tuple[int, int]:
    """[begin, end) of the poses of `view` that device `dev` evaluates in xrc_obj_fn_multi (the library's own
    partition of the camera-major (view, pose) list; host only)."""
    a, n = C.c_uint32(0), C.c_uint32(0)
    _lib.check(_lib.load().xrc_obj_fn_multi_share(n_dev, n_views, n_poses, dev, view, C.byref(a), C.byref(n)))
    return int(a.value), int(a.value + n.value)


class MultiDeviceObjFn:
    """The objective spread over several GPUs of one box from ONE host thread (xrc_obj_fn_multi): one
    Intensity2D3DObjFn replica per device (volume, cameras and fixed images replicated), the camera-major
    (view, pose) projection list cut into contiguous balanced chunks that may straddle views (a
    single-view population is split over the devices, the views of a small multi-view population land on
    different devices), all devices enqueued before any is waited for, only the scalars gathered.
    This is what a single-threaded C++ caller (the reference's optimiser loop) uses; multi-process jobs use
    ShardedObjFn.  `devices` may name the same device more than once (two contexts / streams on one GPU)."""

    def __init__(self, devices: Sequence[int], vol: Volume, cams: Sequence[CameraModel], fixed_imgs: Sequence[np.ndarray],
                 max_pop: int = 100, **kw):
        self.devices = [int(d) for d in devices]
        n_dev = len(self.devices)
        if n_dev == 0:
            raise _lib.XregError("need at least one device")
        self.n_views = len(cams)
        self.max_pop = int(max_pop)
        # a device's chunk of ceil(views * pop / n_dev) projections may lie in a single view
        per_dev = min(self.max_pop, (self.max_pop * self.n_views + n_dev - 1) // n_dev)
        self.ctxs = [Context(d) for d in self.devices]
        self.replicas = [Intensity2D3DObjFn(c, vol, cams, fixed_imgs, max_pop=per_dev, **kw) for c in self.ctxs]
        self._lib = _lib.load()
        self._rc_arr = (C.c_void_p * n_dev)(*[r.rc.handle for r in self.replicas])
        self._sm_arr = (C.c_void_p * (n_dev * self.n_views))(*[sm.handle for r in self.replicas for sm in r.sims])

    def __call__(self, poses: np.ndarray) -> np.ndarray:
        p12 = to12(poses) if np.asarray(poses).ndim == 3 else np.ascontiguousarray(poses, dtype=f32).reshape(-1, 12)
        n = p12.shape[0]
        if n > self.max_pop:
            raise _lib.XregError("population larger than the allocated capacity")
        out = np.empty(n, dtype=f32)
        if n == 0:
            return out
        for r in self.replicas:
            r.rc._flush_params()
        self.per_view = np.empty((self.n_views, n), dtype=f32)
        FP = C.POINTER(C.c_float)
        _lib.check(self._lib.xrc_obj_fn_multi(len(self.replicas), self._rc_arr, self._sm_arr, 0, self.n_views, n,
                                              p12.ctypes.data_as(FP), out.ctypes.data_as(FP),
                                              self.per_view.ctypes.data_as(FP)))
        for r in self.replicas:
            r.rc._poses_dirty = False
            r._cur_pop = -1  # the library re-sized the replica; re-bind on the next direct call
        return out

    def close(self) -> None:
        for r in self.replicas:   # objects first, then the contexts they live on
            r.close()
        self.replicas = []
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unit_chunks(n_units: int, world_size: int) -> List[Tuple[int, int]]:
    """[begin, end) of every rank's chunk of the camera-major (view, pose) unit list: shard_bounds() over units
    (the partition of xrc_obj_fn_multi / xrc_obj_fn_multi_share)."""
    return shard_bounds(n_units, world_size)


def combine_mean(per_view: np.ndarray) -> np.ndarray:
    """ImgSimMetric2DCombineMean::compute (xregImgSimMetric2DCombine.cpp:67-86): sequential f32 sum over views / n_views."""
    pv = np.asarray(per_view, dtype=f32)
    acc = np.zeros(pv.shape[1], dtype=f32)
    for v in range(pv.shape[0]):
        acc = (acc + pv[v]).astype(f32)
    return (acc / f32(pv.shape[0])).astype(f32)


class ShardedViewObjFn:
    """(view, pose)-sharded objective over a torch.distributed process group (config C4: "views sharded across GPUs").

    The camera-major list of n_views x n projections is cut into world_size contiguous balanced chunks that may
    straddle views; local_units_fn(poses, first_unit, n_units) evaluates this rank's chunk (on a GPU rank:
    Intensity2D3DObjFn.eval_units) and returns its per-view values; the ranks all-gather the scalars and every rank
    averages over views.  A population of one pose and three views keeps three ranks busy instead of one."""

    def __init__(self, local_units_fn: Callable[[np.ndarray, int, int], np.ndarray], n_views: int, rank: int = 0,
                 world_size: int = 1, device: str = "cpu", group=None):
        self.local_units_fn = local_units_fn
        self.n_views = int(n_views)
        self.rank, self.world_size = int(rank), int(world_size)
        self.device = device
        self.group = group
        self.per_view: Optional[np.ndarray] = None

    def __call__(self, poses: np.ndarray) -> np.ndarray:
        poses = np.asarray(poses)
        n = poses.shape[0]
        if n == 0:
            return np.zeros(0, dtype=f32)
        bounds = unit_chunks(self.n_views * n, self.world_size)
        b, e = bounds[self.rank]
        local = np.asarray(self.local_units_fn(poses, b, e - b), dtype=f32) if e > b else np.zeros(0, dtype=f32)
        if self.world_size == 1:
            flat = local
        else:
            import torch
            import torch.distributed as dist

            width = max(hi - lo for lo, hi in bounds)
            send = torch.zeros(width, dtype=torch.float32, device=self.device)
            if e > b:
                send[: e - b] = torch.from_numpy(local).to(self.device)
            parts = [torch.empty(width, dtype=torch.float32, device=self.device) for _ in range(self.world_size)]
            dist.all_gather(parts, send, group=self.group)
            recv = torch.stack(parts).cpu().numpy()
            flat = np.concatenate([recv[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)]).astype(f32)
        self.per_view = flat.reshape(self.n_views, n)
        return combine_mean(self.per_view)


class ShardedObjFn:
    """Pose-sharded objective over a torch.distributed process group.

    local_fn evaluates this rank's slice; only the similarity scalars are gathered
    (all_gather of <= ceil(n/world) floats per rank: NCCL over NVLink on GPUs, gloo in
    the CPU tests).  Every rank returns the full (n,) vector in population order."""

    def __init__(self, local_fn: Callable[[np.ndarray], np.ndarray], rank: int = 0, world_size: int = 1,
                 device: str = "cpu", group=None):
        self.local_fn = local_fn
        self.rank, self.world_size = int(rank), int(world_size)
        self.device = device
        self.group = group

    def __call__(self, poses: np.ndarray) -> np.ndarray:
        poses = np.asarray(poses)
        n = poses.shape[0]
        bounds = shard_bounds(n, self.world_size)
        b, e = bounds[self.rank]
        local = np.asarray(self.local_fn(poses[b:e]), dtype=f32)
        if self.world_size == 1:
            return local
        import torch
        import torch.distributed as dist

        width = max(hi - lo for lo, hi in bounds)
        send = torch.zeros(width, dtype=torch.float32, device=self.device)
        if e > b:
            send[: e - b] = torch.from_numpy(local).to(self.device)
        parts = [torch.empty(width, dtype=torch.float32, device=self.device) for _ in range(self.world_size)]
        dist.all_gather(parts, send, group=self.group)
        recv = torch.stack(parts).cpu().numpy()
        return np.concatenate([recv[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)]).astype(f32)


def view_segments(b: int, e: int, n_views: int, n: int) -> List[Tuple[int, int, int]]:
    """The runs (view, first pose, count) that the units [b, e) of the camera-major list u = view * n + pose cover."""
    out = []
    for v in range(n_views):
        lo, hi = max(b, v * n), min(e, (v + 1) * n)
        if hi > lo:
            out.append((v, lo - v * n, hi - lo))
    return out


class ShardedDeviceObjFn:
    """The (view, pose)-sharded objective of a torch.distributed job with one rank per GPU, gathered ON the device:
    the population of `n` poses x n_views views is the camera-major unit list of the reference, cut into
    world_size contiguous balanced chunks (100 poses on 8 ranks: 13 13 13 13 12 12 12 12); every rank ray casts and
    scores only its chunk (xrc_obj_fn_units_enqueue: poses H2D from the library's pinned staging, kernels, no host
    synchronisation), the per-view scalars are all-gathered straight out of the metrics' device result vectors (NCCL
    over NVLink, `width` floats per rank), copied to pinned host memory, and the stream is synchronised ONCE.  Every
    rank returns the full (n,) vector; values are bitwise those of the single-GPU objective (a pose's value does not
    depend on its batch).  The volume and fixed images are replicated; there is no other data-path collective."""

    def __init__(self, fn: Intensity2D3DObjFn, rank: int, world_size: int, group=None, mode: str = "poses"):
        """mode "poses": every rank ray casts and scores its chunk of the (view, pose) list.  mode "tiles": every rank
        ray casts its detector tiles of ALL projections and stores them into their owners' buffers over NVLink (CUDA IPC
        peer mappings, include/xreg_cuda.h "Tile-sharded objective"), then scores the chunk it owns; `fn` must be
        allocated for the whole population on every rank; the barrier and the gather of the scalars are the library's
        own kernel over the same mappings (xrc_obj_fn_tiles).  mode "tiles-nccl": the same with an NCCL all-reduce as
        the barrier and an NCCL all-gather (comparison).  Same values in every mode, bit for bit."""
        import torch

        self.fn, self.rank, self.world_size, self.group = fn, int(rank), int(world_size), group
        self.mode = "tiles" if mode == "tiles-nccl" else mode
        self.nccl_exchange = mode != "tiles"      # "tiles": barrier + gather by the library's kernel over the peer mappings
        if mode not in ("poses", "tiles", "tiles-nccl"):
            raise _lib.XregError("ShardedDeviceObjFn: mode must be 'poses', 'tiles' or 'tiles-nccl'")
        if self.mode == "tiles":
            import torch.distributed as dist

            handles = [None] * self.world_size
            if self.world_size > 1:
                dist.all_gather_object(handles, fn.rc.peer_export(), group=group)
            else:
                handles = [fn.rc.peer_export()]
            fn.rc.peer_attach(self.world_size, self.rank, handles)
        self.n_views = fn.n_views
        self.device = torch.device("cuda", fn.ctx.device)
        self.stream = torch.cuda.ExternalStream(fn.ctx.stream, device=self.device)
        cap = fn.max_pop * self.n_views
        self._send = torch.zeros(cap, dtype=torch.float32, device=self.device)
        self._recv = torch.zeros(cap * self.world_size, dtype=torch.float32, device=self.device)
        self._host = torch.zeros(cap * self.world_size, dtype=torch.float32).pin_memory()
        self._sims = [device_vector(sm.device_sims(), fn.max_pop, self.device) for sm in fn.sims]
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.per_view: Optional[np.ndarray] = None

    def balance(self, poses: np.ndarray, rounds: int = 3, reps: int = 3) -> List[int]:
        """mode "tiles" (collective: every rank calls it with the same population): let the clock correct the tile
        plan.  Per round every rank times its ray-casting kernel on `poses` (CUDA events on the library's stream, best
        of `reps`), the times are all-gathered and xrc_rc_plan_tiles_timed re-cuts the ranges; every rank derives the
        same plan.  Returns the final tile bounds.  Results never depend on the plan (a pixel is the same whichever GPU
        computes it); worth calling once per registration level, where the pose distribution moves."""
        import torch
        import torch.distributed as dist

        if self.mode != "tiles":
            return []
        plan: List[int] = []
        t = torch.zeros(self.world_size, dtype=torch.float32, device=self.device)
        mine = torch.zeros(1, dtype=torch.float32, device=self.device)
        for _ in range(max(int(rounds), 0)):
            best = float("inf")
            with torch.cuda.stream(self.stream):
                self.fn.enqueue_tiles_drr(poses)      # the first launch (re)builds the plan if there is none
                for _r in range(max(int(reps), 1)):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if self.world_size > 1:
                        dist.all_reduce(self._flag, group=self.group)   # start together: peers' stores share NVLink
                    e0.record(self.stream)
                    self.fn.rc.compute_tiles()
                    e1.record(self.stream)
                    e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                mine.fill_(best)
                if self.world_size > 1:
                    dist.all_gather_into_tensor(t, mine, group=self.group)
                else:
                    t.copy_(mine)
            self.stream.synchronize()
            plan = self.fn.rc.plan_tiles_timed([float(v) for v in t.cpu().tolist()])
        if self.world_size > 1:
            dist.barrier(group=self.group)
        self.last_balance_ms = [float(v) for v in t.cpu().tolist()]
        return plan

    def enqueue(self, poses: np.ndarray):
        """Everything but the final synchronise: returns (bounds, width) for collect()."""
        import torch
        import torch.distributed as dist

        n = np.asarray(poses).shape[0]
        bounds = unit_chunks(self.n_views * n, self.world_size)
        width = max(hi - lo for lo, hi in bounds)
        b, e = bounds[self.rank]
        with torch.cuda.stream(self.stream):
            if self.mode == "tiles":
                self.fn.enqueue_tiles_drr(poses)
                if self.world_size > 1:
                    # barrier ordered on the streams: every rank's tiles have landed in their owners' buffers (the
                    # all-gather below is the barrier that protects those buffers from the NEXT call's stores)
                    dist.all_reduce(self._flag, group=self.group)
                self.fn.enqueue_units_metrics(n, b, e - b)
            else:
                self.fn.enqueue_units(poses, b, e - b)
            segs = view_segments(b, e, self.n_views, n)
            if len(segs) == 1 and width <= self.fn.max_pop:
                send = self._sims[segs[0][0]][:width]          # zero copy: the metric's own result vector
            else:
                send, off = self._send[:width], 0
                for v, _, cnt in segs:
                    send[off:off + cnt].copy_(self._sims[v][:cnt])
                    off += cnt
            recv = self._recv[: width * self.world_size]
            if self.world_size > 1:
                dist.all_gather_into_tensor(recv, send, group=self.group)
            else:
                recv.copy_(send)
            self._host[: width * self.world_size].copy_(recv, non_blocking=True)
        return bounds, width

    def collect(self, bounds, width: int, n: int) -> np.ndarray:
        self.stream.synchronize()
        h = self._host.numpy()
        flat = np.concatenate([h[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]).astype(f32)
        self.per_view = flat.reshape(self.n_views, n)
        return combine_mean(self.per_view) if self.n_views > 1 else self.per_view[0].copy()

    def __call__(self, poses: np.ndarray) -> np.ndarray:
        n = np.asarray(poses).shape[0]
        if n == 0:
            return np.zeros(0, dtype=f32)
        if self.mode == "tiles" and not self.nccl_exchange:
            out = self.fn.eval_tiles(poses)        # one library call, no NCCL on the path
            self.per_view = self.fn.per_view
            return out
        bounds, width = self.enqueue(poses)
        return self.collect(bounds, width, n)


def device_vector(ptr: int, n: int, device):
    """Zero-copy torch view of n device floats at `ptr` (memory owned by the library)."""
    import torch

    class _Ptr:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Ptr(), device=device)
