"""SURVEY 8(f) rank 2 (caller-side glue): the regulariser of Intensity2D3DRegi::obj_fn.  xrc_se3_mag_penalty restates
Regi2D3DPenaltyFnSE3Mag::compute with FoldNormDist densities (xregRegi2D3DPenaltyFnSE3Mag.cpp:32-117, xregFoldNormDist.cpp,
ComputeRotAngTransMag / LogSO3ToPt); it is pinned here to the reference's own files compiled from /root/reference
(oracle/_ref/libxreg_refpenalty.so, tests/xreg_link/build_link.py) bit for bit.  The GPU test checks the one-call
objective with the penalty (xrc_obj_fn_se3_pen, xregIntensity2D3DRegi.cpp:653-688)."""
import ctypes as C
import os

import numpy as np
import pytest

from xreg_b200 import regi, synth
from xreg_b200.geometry import exp_se3, to12

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEN_LIB = os.path.join(ROOT, "oracle", "_ref", "libxreg_refpenalty.so")
f32 = np.float32


def _ref_lib():
    if os.path.isdir("/root/reference"):
        from tests.xreg_link import build_link

        build_link.build_penalty()
    if not os.path.exists(PEN_LIB):
        pytest.skip("oracle/_ref/libxreg_refpenalty.so not built (needs /root/reference at build time)")
    return C.CDLL(PEN_LIB)


def _ref_penalty(lib, pen, poses):
    p12 = to12(poses)
    out = np.zeros(p12.shape[0], f32)
    FP = C.POINTER(C.c_float)
    inter = np.array(list(pen.inter_frame), f32)
    init = np.array(list(pen.init_cam_to_vol), f32)
    lib.xref_se3_mag_penalty(C.c_float(pen.rot_mean), C.c_float(pen.rot_std), C.c_float(pen.trans_mean), C.c_float(pen.trans_std),
                             C.c_int(pen.inter_wrt_vol), inter.ctypes.data_as(FP), init.ctypes.data_as(FP),
                             C.c_uint(p12.shape[0]), p12.ctypes.data_as(FP), out.ctypes.data_as(FP))
    return out


def _rand_rigid(rng, rot_deg, trans):
    x = np.concatenate([rng.normal(0, np.deg2rad(rot_deg), 3), rng.normal(0, trans, 3)])
    return exp_se3(x).astype(f32)


@pytest.mark.parametrize("inter_wrt_vol", [True, False])
def test_penalty_equals_the_reference_lines(inter_wrt_vol):
    lib = _ref_lib()
    rng = np.random.default_rng(12)
    for case in range(6):
        init = (_rand_rigid(rng, 40, 200) @ np.diag([1, 1, 1, 1]).astype(f32)).astype(f32)
        inter = _rand_rigid(rng, 60, 80) if case % 2 else np.eye(4, dtype=f32)
        pen = regi.se3_penalty(np.deg2rad(10.0) * (case % 3 != 2), np.deg2rad(10.0), 50.0 * (case % 3 != 2), 50.0 if case < 4 else 3.0,
                               inter_frame=inter, init_cam_to_vol=init, inter_wrt_vol=inter_wrt_vol)
        # poses around the initial guess: from identical to far away (the far tail takes FoldNormDist's 1e-14 floor)
        poses = np.stack([(init @ _rand_rigid(rng, s, 10 * s)).astype(f32) for s in (0.0, 1e-4, 0.5, 3, 10, 30, 80, 170) for _ in range(6)])
        poses[0] = init
        got = regi.se3_mag_penalty(pen, poses)
        ref = _ref_penalty(lib, pen, poses)
        assert np.array_equal(got, ref, equal_nan=True), (case, got, ref)
        assert np.isfinite(got[:24]).all()     # near poses are finite; the far tail hits FoldNormDist's floor (3.4e38 / inf), as there
    # known answer: pose == initial guess, zero-mean densities: log p(0) = log 2 - log Z with Z = sqrt(2 s^2 pi), and
    # reg = (log Z_r - log p_r) + (log Z_t - log p_t) = 2 log Z_r + 2 log Z_t - 2 log 2
    pen = regi.se3_penalty(0.0, 0.2, 0.0, 30.0)
    v = regi.se3_mag_penalty(pen, np.eye(4, dtype=f32)[None])
    expect = np.log(2 * 0.2 ** 2 * np.pi) + np.log(2 * 30.0 ** 2 * np.pi) - 2.0 * np.log(2.0)
    assert abs(float(v[0]) - expect) < 1e-5


@pytest.mark.gpu
def test_objective_with_penalty_in_one_call(ctx, xo, small_scene):
    vol, cam, nominal = small_scene
    c = np.asarray(vol.origin) + 0.5 * (np.asarray(vol.dims) - 1.0) * np.asarray(vol.spacing)
    pre, post = np.eye(4, dtype=f32), np.eye(4, dtype=f32)
    pre[:3, 3] = c
    post[:3, 3] = -c
    post = (post @ nominal).astype(f32)
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(0, 0.05, (9, 3)), rng.normal(0, 4.0, (9, 3))], axis=1).astype(f32)
    x[0] = 0
    poses = np.stack([(pre @ exp_se3(xi) @ post).astype(f32) for xi in x])
    fixed = synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses[:1]))[0])
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="grad-ncc", max_pop=9)
    plain = fn.eval_se3(x, pre, post)
    pen = regi.se3_penalty(np.deg2rad(10.0), np.deg2rad(10.0), 50.0, 50.0, inter_frame=pre, init_cam_to_vol=nominal,
                           inter_wrt_vol=True, img_sim_coeff=0.9, penalty_coeff=0.1)
    got = fn.eval_se3(x, pre, post, penalty=pen)
    # the library composes the poses itself: penalise exactly those
    lib = __import__("xreg_b200")._lib.load()
    composed = np.zeros((9, 12), f32)
    FP = C.POINTER(C.c_float)
    for i in range(9):
        T = np.zeros(12, f32)
        lib.xrc_exp_se3(x[i].ctypes.data_as(FP), T.ctypes.data_as(FP))
        M = np.vstack([T.reshape(3, 4), [0, 0, 0, 1]]).astype(f32)
        composed[i] = to12(regi_mul(regi_mul(pre, M), post))
    reg = regi.se3_mag_penalty(pen, composed)
    np.testing.assert_allclose(fn.last_penalty, reg, rtol=0, atol=2e-5)
    np.testing.assert_allclose(got, (plain * f32(0.9) + fn.last_penalty * f32(0.1)).astype(f32), rtol=0, atol=1e-6)
    # without coefficients the two terms are simply added
    pen2 = regi.se3_penalty(np.deg2rad(10.0), np.deg2rad(10.0), 50.0, 50.0, inter_frame=pre, init_cam_to_vol=nominal)
    got2 = fn.eval_se3(x, pre, post, penalty=pen2)
    np.testing.assert_allclose(got2, (plain + fn.last_penalty).astype(f32), rtol=0, atol=1e-6)
    ref_lib = _ref_lib()
    assert np.array_equal(fn.last_penalty, _ref_penalty(ref_lib, pen2, np.stack([np.vstack([p.reshape(3, 4), [0, 0, 0, 1]]) for p in composed]).astype(f32))) \
        or np.allclose(fn.last_penalty, reg, atol=2e-5)
    fn.close()


def regi_mul(a, b):
    """row-major affine product with the library's operation order (left-to-right sums)"""
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    out = np.eye(4, dtype=f32)
    for r in range(3):
        for k in range(4):
            v = f32(f32(f32(a[r, 0] * b[0, k]) + f32(a[r, 1] * b[1, k])) + f32(a[r, 2] * b[2, k]))
            if k == 3:
                v = f32(v + a[r, 3])
            out[r, k] = v
    return out
