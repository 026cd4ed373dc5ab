"""ctypes binding of libxreg_cuda.so (the C ABI in include/xreg_cuda.h).

There is no fallback of any kind: if the shared library is missing the import
fails with instructions to build it, and every compute call fails loudly when
no sm_100-class GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxreg_cuda.so")

XRC_OK = 0
XRC_ERR_INVALID = 1
XRC_ERR_UNSUPPORTED = 2
XRC_ERR_CUDA = 3
XRC_ERR_NOMEM = 4

INTERP_LINEAR, INTERP_NN, INTERP_SINC, INTERP_BSPLINE = 0, 1, 2, 3
STORE_REPLACE, STORE_ACCUM = 0, 1
KERNEL_SUM, KERNEL_MAX = 0, 1
SM_NCC, SM_GRAD_NCC, SM_PATCH_NCC, SM_PATCH_GRAD_NCC, SM_SSD = 0, 1, 2, 3, 4
LAYOUT_DEFAULT, LAYOUT_LINEAR, LAYOUT_QUAD, LAYOUT_TEX_QUAD, LAYOUT_OCT, LAYOUT_TEX = -1, 0, 1, 2, 3, 4
LAYOUT_NAMES = {"default": -1, "linear": 0, "quad": 1, "tex_quad": 2, "oct": 3, "tex": 4, "pax": 5}


class XrcCam(C.Structure):
    """xrc_cam (include/xreg_cuda.h)"""

    _fields_ = [
        ("rows", C.c_uint32),
        ("cols", C.c_uint32),
        ("intrins_inv", C.c_float * 9),
        ("extrins_inv", C.c_float * 12),
        ("pinhole", C.c_float * 3),
        ("focal_len", C.c_float),
        ("frame_type", C.c_int32),
    ]


class XrcSe3Penalty(C.Structure):
    """xrc_se3_penalty (include/xreg_cuda.h)"""

    _fields_ = [
        ("rot_mean", C.c_float), ("rot_std", C.c_float), ("trans_mean", C.c_float), ("trans_std", C.c_float),
        ("use_coeffs", C.c_int32), ("img_sim_coeff", C.c_float), ("penalty_coeff", C.c_float),
        ("inter_wrt_vol", C.c_int32), ("inter_frame", C.c_float * 12), ("init_cam_to_vol", C.c_float * 12),
    ]


class XregError(RuntimeError):
    """StringMessageException / AssertFailedException analogue
    (lib/common/xregExceptionUtils.h:36-63, lib/common/xregAssert.h:30-41)."""


class UnsupportedOperationException(XregError):
    """RayCaster::UnsupportedOperationException (lib/ray_cast/xregRayCastInterface.h:59)"""


class XregCudaError(XregError):
    """CUDA failure or no usable GPU (there is no CPU fallback)."""


_VP = C.c_void_p
_U32 = C.c_uint32
_U64 = C.c_uint64
_FP = C.POINTER(C.c_float)
_U8P = C.POINTER(C.c_uint8)
_U32P = C.POINTER(C.c_uint32)

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "xrc_last_error": [],
    "xrc_version": [],
    "xrc_launch_count": [],
    "xrc_ctx_create": [C.c_int, C.POINTER(_VP)],
    "xrc_ctx_create_on_stream": [C.c_int, _VP, C.POINTER(_VP)],
    "xrc_ctx_destroy": [_VP],
    "xrc_ctx_synchronize": [_VP],
    "xrc_ctx_device": [_VP, C.POINTER(C.c_int)],
    "xrc_ctx_stream": [_VP, C.POINTER(_VP)],
    "xrc_rc_create": [_VP, C.POINTER(_VP)],
    "xrc_rc_destroy": [_VP],
    "xrc_rc_set_layout": [_VP, C.c_int],
    "xrc_rc_set_cta_order": [_VP, C.c_int],
    "xrc_rc_set_volumes": [_VP, _U32, C.POINTER(_FP), C.POINTER(_U64 * 3), C.POINTER(C.c_float * 12)],
    "xrc_rc_set_volumes_hu": [_VP, _U32, C.POINTER(_FP), C.POINTER(_U64 * 3), C.POINTER(C.c_float * 12), C.c_float],
    "xrc_rc_set_volumes_device": [_VP, _U32, C.POINTER(_VP), C.POINTER(_U64 * 3), C.POINTER(C.c_float * 12)],
    "xrc_rc_set_cameras": [_VP, _U32, C.POINTER(XrcCam)],
    "xrc_rc_allocate": [_VP, _U32],
    "xrc_rc_set_num_projs": [_VP, _U32],
    "xrc_rc_num_projs": [_VP, _U32P],
    "xrc_rc_max_projs_possible": [_VP, C.POINTER(_U64)],
    "xrc_rc_set_poses": [_VP, _U32, _FP, _U32P],
    "xrc_rc_distribute_poses": [_VP, _U32, _FP],
    "xrc_rc_set_poses_device": [_VP, _U32, _VP, _VP],
    "xrc_rc_set_poses_device_mirrored": [_VP, _U32, _VP, _VP, _FP, _U32P],
    "xrc_rc_set_params": [_VP, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float],
    "xrc_rc_set_bg_projs": [_VP, C.POINTER(_FP), C.c_int],
    "xrc_rc_compute": [_VP, _U32],
    "xrc_rc_device_buf": [_VP, C.POINTER(_VP)],
    "xrc_rc_read_projs": [_VP, _U32, _U32, _FP],
    "xrc_rc_use_other_proj_buf": [_VP, _VP],
    "xrc_rc_ray_info": [_VP, _U32, _U8P, _U32P, C.POINTER(_U64)],
    "xrc_rc_volume_bytes": [_VP, C.POINTER(_U64)],
    "xrc_rc_volume_layout": [_VP, _U32, C.POINTER(C.c_int)],
    "xrc_rc_set_skip_empty": [_VP, C.c_int],
    "xrc_rc_fetched_samples": [_VP, _U32, C.POINTER(_U64)],
    "xrc_sm_create": [_VP, C.c_int, C.POINTER(_VP)],
    "xrc_sm_destroy": [_VP],
    "xrc_sm_set_fixed": [_VP, _FP, _U32, _U32],
    "xrc_sm_set_mask": [_VP, _U8P],
    "xrc_sm_set_grad_params": [_VP, _U32],
    "xrc_sm_set_patch_params": [_VP, _U32, _U32, C.c_int, C.c_int, C.c_int, _FP, _U64],
    "xrc_sm_set_patch_subset": [_VP, C.POINTER(_U64), _U64],
    "xrc_sm_set_combine_mode": [_VP, C.c_int],
    "xrc_seqsum_f32": [_VP, _FP, _U32, _U64, C.c_int, _FP],
    "xrc_sm_bind_ray_caster": [_VP, _VP, _U32],
    "xrc_sm_bind_host": [_VP, _FP, _U32],
    "xrc_sm_bind_device": [_VP, _VP, _U32],
    "xrc_sm_allocate": [_VP, _U32],
    "xrc_sm_set_num_imgs": [_VP, _U32],
    "xrc_sm_compute": [_VP],
    "xrc_sm_read_sims": [_VP, _FP, _U32],
    "xrc_sm_device_sims": [_VP, C.POINTER(_VP)],
    "xrc_sm_read_grads": [_VP, _U32, _FP, _FP],
    "xrc_eval_batch": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP],
    "xrc_eval_batch_async": [_VP, _U32, C.POINTER(_VP), _U32],
    "xrc_obj_fn": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _FP, _FP],
    "xrc_obj_fn_objects": [_VP, _U32, _U32P, C.POINTER(_VP), _U32, _U32, _FP, C.c_int, _FP, _FP],
    "xrc_obj_fn_multi": [_U32, C.POINTER(_VP), C.POINTER(_VP), _U32, _U32, _U32, _FP, _FP, _FP],
    "xrc_obj_fn_se3": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _FP, _FP, _FP, _FP],
    "xrc_obj_fn_units": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _U32, _U32, _FP],
    "xrc_obj_fn_units_enqueue": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _U32, _U32],
    "xrc_rc_compute_depth": [_VP, _U32, C.c_float, _U32],
    "xrc_log_remap": [_VP, _FP, _U32, _U32, C.c_int, C.c_int, C.c_float, _FP, _FP],
    "xrc_downsample_size": [_U32, _U32, C.c_double, _U32P, _U32P],
    "xrc_downsample_image": [_VP, _FP, _U32, _U32, C.c_double, C.c_double, _FP],
    "xrc_rc_peer_export": [_VP, C.POINTER(C.c_uint8)],
    "xrc_rc_peer_attach": [_VP, _U32, _U32, C.POINTER(C.c_uint8)],
    "xrc_rc_peer_detach": [_VP],
    "xrc_rc_compute_tiles": [_VP, _U32],
    "xrc_rc_plan_tiles": [_VP, _U32],
    "xrc_rc_plan_tiles_timed": [_VP, _U32, _FP],
    "xrc_rc_tile_plan": [_VP, _U32P],
    "xrc_rc_tile_samples": [_VP, _U32, C.POINTER(_U64), C.POINTER(_U64)],
    "xrc_obj_fn_tiles_enqueue_drr": [_VP, _U32, _U32, _U32, _FP],
    "xrc_obj_fn_units_enqueue_metrics": [_VP, C.POINTER(_VP), _U32, _U32, _U32, _U32],
    "xrc_rc_peer_barrier": [_VP],
    "xrc_obj_fn_tiles_enqueue_gather": [_VP, C.POINTER(_VP), _U32, _U32],
    "xrc_obj_fn_tiles_finish": [_VP, _U32, _U32, _FP, _FP],
    "xrc_obj_fn_tiles": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _FP, _FP],
    "xrc_obj_fn_multi_share": [_U32, _U32, _U32, _U32, _U32, _U32P, _U32P],
    "xrc_exp_se3": [_FP, _FP],
    "xrc_se3_mag_penalty": [C.POINTER(XrcSe3Penalty), _U32, _FP, _FP],
    "xrc_obj_fn_se3_pen": [_VP, _U32, C.POINTER(_VP), _U32, _U32, _FP, _FP, _FP, C.POINTER(XrcSe3Penalty), _FP, _FP, _FP],
}
_RESTYPES = {"xrc_last_error": C.c_char_p, "xrc_launch_count": C.c_uint64, "xrc_exp_se3": None}

_lib = None


def load():
    """Load libxreg_cuda.so and declare every entry point.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # a fresh checkout (built artefacts are not in git): compile the product once, or fail loudly
        csrc = os.path.join(os.path.dirname(LIB_PATH), "csrc")
        try:
            import subprocess

            subprocess.run(["make", "-C", csrc], check=True, capture_output=True)
        except Exception as e:  # no nvcc, compile error, ...
            raise ImportError(
                "%s not found and `make -C xreg_b200/csrc` failed (%s). xreg_b200 has no CPU or PyTorch fallback."
                % (LIB_PATH, e))
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found after the build. xreg_b200 has no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int) -> None:
    """Turn an xrc_status into the reference's exception types."""
    if status == XRC_OK:
        return
    msg = load().xrc_last_error()
    msg = msg.decode("utf-8", "replace") if msg else "unknown error"
    if status == XRC_ERR_UNSUPPORTED:
        raise UnsupportedOperationException(msg)
    if status in (XRC_ERR_CUDA, XRC_ERR_NOMEM):
        raise XregCudaError(msg)
    raise XregError(msg)


def launch_count() -> int:
    return int(load().xrc_launch_count())
