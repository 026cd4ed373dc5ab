"""Config C3 in miniature: the library as the objective of a real multi-resolution registration loop
(CMA-ES populations at the coarse level, population-1 refinement at the finer one; scripts/register_c3.py),
with the oracle re-evaluating the poses the optimiser visited.

Reference call structure: MultiLevelMultiObjRegi::run (xregMultiObjMultiLevel2D3DRegi.cpp:164-539),
Intensity2D3DRegiCMAES::run (xregIntensity2D3DRegiCMAES.cpp:75-261), Intensity2D3DRegi::obj_fn
(xregIntensity2D3DRegi.cpp:571-696).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import register_c3 as R  # noqa: E402
from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import exp_se3, to12  # noqa: E402

f32 = np.float32


def test_cmaes_stand_in_minimises_an_ill_conditioned_quadratic():
    """Host-only: the optimiser stand-in used by the C3 script and the GPU test below does its job."""
    A = np.diag([1.0, 10.0, 100.0, 1.0, 1000.0, 3.0])
    x_opt = np.array([0.3, -0.2, 0.1, 4.0, -2.0, 7.0])
    es = R.CMAES(np.zeros(6), [0.5, 0.5, 0.5, 5.0, 5.0, 10.0], popsize=24, seed=1)
    for _ in range(150):
        X = es.ask()
        es.tell(np.einsum("ni,ij,nj->n", X - x_opt, A, X - x_opt))
    assert np.max(np.abs(es.mean - x_opt)) < 1e-4


def _scene():
    vol = synth.make_volume(96, 96, 72, spacing=(1.6, 1.6, 2.0))
    truth = synth.nominal_pose(vol).astype(np.float64)
    centre = np.asarray(vol.origin) + 0.5 * (np.array(vol.dims) - 1.0) * np.asarray(vol.spacing)
    C4, Ci4 = np.eye(4), np.eye(4)
    C4[:3, 3], Ci4[:3, 3] = centre, -centre
    off = np.array([np.deg2rad(4.0), np.deg2rad(-3.0), np.deg2rad(5.0), 8.0, -6.0, 15.0])
    return vol, truth, centre, C4 @ exp_se3(off) @ Ci4 @ truth


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["grad-ncc", "patch-grad-ncc"])
def test_multi_resolution_registration_recovers_the_pose(xo, metric):
    vol, truth, centre, init = _scene()
    make_objective, render_fixed, ctx = R.gpu_factories(vol, metric, 50, compose_on_host=True)
    visited = []   # (detector, fixed image, poses, CUDA values) of a few objective calls, re-evaluated by the oracle

    def make_objective_logged(det, fixed):
        fn, close = make_objective(det, fixed)

        def f(X, pre, post):
            v = fn(X, pre, post)
            if len(visited) < 64 and (len(X) > 1 or len(visited) % 2 == 0):
                visited.append((det, fixed, np.array(X[:6], f32), np.asarray(pre), np.asarray(post), np.array(v[:6], f32)))
            return v

        return f, close

    res = R.run_registration(make_objective_logged, render_fixed, centre, truth, init,
                             levels=((48, "cmaes"), (96, "local")), popsize=50, cma_gens=30, local_evals=300, seed=0)
    ctx.close()
    assert res["init_error"]["rot_deg"] > 6.0 and res["init_error"]["trans_mm"] > 15.0
    # single view: rotation and in-plane position are well determined, depth less so
    assert res["final_error"]["rot_deg"] < 0.5, res["final_error"]
    assert res["final_error"]["in_plane_mm"] < 0.5, res["final_error"]
    assert res["final_error"]["depth_mm"] < 3.0, res["final_error"]
    assert res["levels"][0]["pose_evals"] == 50 * 30 + 1 and res["levels"][0]["objective_calls"] == 31

    # the values the optimiser saw are the CPU classes' values (<= 1e-5) at those poses
    worst = 0.0
    for det, fixed, X, pre, post, v in visited[::4]:
        cam = xo.cam_struct(synth.make_camera(det))
        poses = R.host_poses(X, pre, post)
        d = xo.drr(vol.data, vol.idx_to_phys(), [cam], to12(poses))
        if metric == "grad-ncc":
            ref = xo.grad_ncc(fixed, d, gauss_width=5)
        else:
            ref = xo.patch_grad_ncc(fixed, d, xo.patch_opts(radius=synth.patch_radius_for(det)), gauss_width=5)
        worst = max(worst, float(np.max(np.abs(ref - v))))
    assert worst <= 1e-5, worst
