#!/usr/bin/env python
"""Config C3 end to end: a multi-resolution single-view registration that uses the library the way
xReg's pipeline does (MultiLevelMultiObjRegi::run, xregMultiObjMultiLevel2D3DRegi.cpp:164-539):

  level 0: detector 1536/8 = 192^2, CMA-ES, population 100 per objective call  (Intensity2D3DRegiCMAES::run,
           xregIntensity2D3DRegiCMAES.cpp:75-261: every generation hands the whole population to obj_fn once)
  level 1: detector 1536/4 = 384^2, local derivative-free refinement, population 1 per call (the reference
           uses NLopt BOBYQA, xregIntensity2D3DRegiNLOptInterface.cpp:255-288)
  level 2: detector 1536/2 = 768^2, the same

The optimisers are the CALLERS of the hot path and are out of scope of this repo;
the two below are stand-ins written for this script from the published algorithms -- a plain (mu/mu_w, lambda)-CMA-ES
(Hansen's tutorial; the reference links c-cmaes) and scipy's Nelder-Mead in place of BOBYQA (NLopt is not in this image).
What the script measures is the library under a real optimiser loop: objective calls, pose evaluations, wall time per
level (host optimiser work included), and how well the known pose is recovered.

Every objective call goes through xrc_obj_fn_se3 (pose_p = pre * ExpSE3(x_p) * post composed inside the library,
xregIntensity2D3DRegi.cpp:1049-1071 with the SE3OptVarsLieAlg parameterisation, xregSE3OptVars.cpp:128-137).
GPU only.  Prints one JSON line per level and a summary line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import exp_se3, se3_inv  # noqa: E402

# standard deviations of the reference app's CMA-ES stage (pelvis...main.cpp:292): 15, 15, 30 deg, 50, 50, 100 mm
APP_SIGMA = np.array([np.deg2rad(15.0), np.deg2rad(15.0), np.deg2rad(30.0), 50.0, 50.0, 100.0])


class CMAES:
    """Minimal (mu/mu_w, lambda)-CMA-ES with per-coordinate initial standard deviations (ask / tell)."""

    def __init__(self, x0: Sequence[float], sigma0: Sequence[float], popsize: int, seed: int = 0):
        self.n = n = len(x0)
        self.lam = int(popsize)
        self.mean = np.asarray(x0, dtype=np.float64).copy()
        self.scale = np.asarray(sigma0, dtype=np.float64).copy()   # coordinates are optimised in units of sigma0
        self.sigma = 1.0
        self.rng = np.random.default_rng(seed)
        mu = self.lam // 2
        w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
        self.w = w / w.sum()
        self.mu = mu
        self.mueff = 1.0 / np.sum(self.w ** 2)
        self.cc = (4 + self.mueff / n) / (n + 4 + 2 * self.mueff / n)
        self.cs = (self.mueff + 2) / (n + self.mueff + 5)
        self.c1 = 2 / ((n + 1.3) ** 2 + self.mueff)
        self.cmu = min(1 - self.c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((n + 2) ** 2 + self.mueff))
        self.damps = 1 + 2 * max(0.0, np.sqrt((self.mueff - 1) / (n + 1)) - 1) + self.cs
        self.chin = np.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        self.pc = np.zeros(n)
        self.ps = np.zeros(n)
        self.Cm = np.eye(n)
        self.B = np.eye(n)
        self.D = np.ones(n)
        self.gen = 0
        self._y = None

    def ask(self) -> np.ndarray:
        z = self.rng.standard_normal((self.lam, self.n))
        self._y = (z * self.D) @ self.B.T
        return self.mean + self.sigma * self._y * self.scale

    def tell(self, f: np.ndarray) -> None:
        n = self.n
        order = np.argsort(f)[: self.mu]
        ysel = self._y[order]
        yw = self.w @ ysel
        self.mean = self.mean + self.sigma * yw * self.scale
        invsqrt = self.B @ np.diag(1.0 / self.D) @ self.B.T
        self.ps = (1 - self.cs) * self.ps + np.sqrt(self.cs * (2 - self.cs) * self.mueff) * (invsqrt @ yw)
        self.gen += 1
        hsig = (np.linalg.norm(self.ps) / np.sqrt(1 - (1 - self.cs) ** (2 * self.gen)) / self.chin) < (1.4 + 2 / (n + 1))
        self.pc = (1 - self.cc) * self.pc + (np.sqrt(self.cc * (2 - self.cc) * self.mueff) * yw if hsig else 0.0)
        rank_mu = (ysel * self.w[:, None]).T @ ysel
        self.Cm = ((1 - self.c1 - self.cmu) * self.Cm
                   + self.c1 * (np.outer(self.pc, self.pc) + (0.0 if hsig else self.cc * (2 - self.cc)) * self.Cm)
                   + self.cmu * rank_mu)
        self.sigma *= np.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chin - 1))
        self.Cm = 0.5 * (self.Cm + self.Cm.T)
        d2, self.B = np.linalg.eigh(self.Cm)
        self.D = np.sqrt(np.maximum(d2, 1e-20))


def pose_error(T_est: np.ndarray, T_true: np.ndarray, centre: np.ndarray) -> Dict[str, float]:
    """Rotation angle (deg) and displacement of the volume centre (mm, total and along the viewing axis)
    between two cam -> volume-physical transforms, measured in the camera frame."""
    A, Bm = se3_inv(np.asarray(T_est, np.float64)), se3_inv(np.asarray(T_true, np.float64))   # volume -> camera
    dR = A[:3, :3] @ Bm[:3, :3].T
    ang = np.degrees(np.arccos(np.clip(0.5 * (np.trace(dR) - 1.0), -1.0, 1.0)))
    c = np.append(centre, 1.0)
    d = (A @ c - Bm @ c)[:3]
    return {"rot_deg": float(ang), "trans_mm": float(np.linalg.norm(d)), "in_plane_mm": float(np.linalg.norm(d[:2])),
            "depth_mm": float(abs(d[2]))}


def run_registration(make_objective: Callable[[int, np.ndarray], Tuple[Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray], Callable[[], None]]],
                     render_fixed: Callable[[int, np.ndarray], np.ndarray], centre: np.ndarray, truth: np.ndarray,
                     init: np.ndarray, levels: Sequence[Tuple[int, str]], popsize: int = 100, cma_gens: int = 40,
                     sigma0: Optional[np.ndarray] = None, local_evals: int = 400, seed: int = 0,
                     log: Optional[Callable[[dict], None]] = None) -> dict:
    """levels: (detector size, "cmaes" | "local").  make_objective(det, fixed) -> (f(params (n,6), pre, post) -> (n,), close).
    The current estimate is carried from level to level as in the reference pipeline (regi k starts at regi k-1's result)."""
    C4, Ci4 = np.eye(4), np.eye(4)
    C4[:3, 3], Ci4[:3, 3] = centre, -centre
    cur = np.asarray(init, dtype=np.float64)
    sigma0 = APP_SIGMA / 3.0 if sigma0 is None else np.asarray(sigma0, np.float64)
    out_levels: List[dict] = []
    total_evals, total_s = 0, 0.0
    for det, kind in levels:
        fixed = render_fixed(det, truth)
        fn, close = make_objective(det, fixed)
        pre = C4.astype(np.float32)
        post = (Ci4 @ cur).astype(np.float32)       # pose(x) = C exp(x) C^-1 cur, x = 0 is the current estimate
        calls = evals = 0
        fn(np.zeros((1, 6), np.float32), pre, post)  # warm-up (allocation, first launch) outside the timed region
        t0 = time.perf_counter()
        if kind == "cmaes":
            es = CMAES(np.zeros(6), sigma0, popsize, seed=seed)
            best_x, best_f = np.zeros(6), np.inf
            for _ in range(cma_gens):
                X = es.ask()
                f = np.asarray(fn(X.astype(np.float32), pre, post), dtype=np.float64)
                calls += 1
                evals += len(X)
                es.tell(f)
                k = int(np.argmin(f))
                if f[k] < best_f:
                    best_f, best_x = float(f[k]), X[k].copy()
            # the reference takes the distribution mean ("xmean") as the result (xregIntensity2D3DRegiCMAES.cpp:236-248)
            x_fin = es.mean
            f_fin = float(fn(x_fin[None].astype(np.float32), pre, post)[0])
            calls += 1
            evals += 1
            if best_f < f_fin:
                x_fin, f_fin = best_x, best_f
        else:
            unit = np.array([np.deg2rad(1.0)] * 3 + [1.0, 1.0, 2.0])   # optimise in ~1 deg / 1 mm units

            def f1(u):
                nonlocal calls, evals
                calls += 1
                evals += 1
                return float(fn((u * unit)[None].astype(np.float32), pre, post)[0])

            simplex0 = np.vstack([np.zeros(6)] + [np.eye(6)[i] * (1.0 if det <= 384 else 0.5) for i in range(6)])
            res = minimize(f1, np.zeros(6), method="Nelder-Mead",
                           options={"maxfev": local_evals, "xatol": 1e-3, "fatol": 1e-9, "initial_simplex": simplex0})
            x_fin, f_fin = res.x * unit, float(res.fun)
        dt = time.perf_counter() - t0
        close()
        cur = C4 @ exp_se3(x_fin) @ Ci4 @ cur
        rec = {"level_det": det, "optimiser": kind, "objective_calls": calls, "pose_evals": evals, "seconds": dt,
               "pose_evals_per_s": evals / dt, "ms_per_call": 1e3 * dt / max(calls, 1), "final_sim": f_fin,
               **{"err_" + k: v for k, v in pose_error(cur, truth, centre).items()}}
        out_levels.append(rec)
        total_evals += evals
        total_s += dt
        if log:
            log(rec)
    return {"levels": out_levels, "pose_evals": total_evals, "seconds": total_s, "final_pose": cur,
            "init_error": pose_error(init, truth, centre), "final_error": pose_error(cur, truth, centre)}


def host_poses(X: np.ndarray, pre: np.ndarray, post: np.ndarray) -> np.ndarray:
    """pre * ExpSE3(x_p) * post in f64, rounded to f32: (n, 4, 4)."""
    pre, post = np.asarray(pre, np.float64), np.asarray(post, np.float64)
    return np.stack([(pre @ exp_se3(x) @ post).astype(np.float32) for x in np.asarray(X, np.float64).reshape(-1, 6)])


def gpu_factories(vol, metric: str, popsize: int, noise: float = 0.01, compose_on_host: bool = False):
    """(make_objective, render_fixed, ctx) over the CUDA library for a single-view acquisition.  compose_on_host: build
    the poses pre * ExpSE3(x) * post here (f64, rounded to f32) and call xrc_obj_fn instead of xrc_obj_fn_se3, so that a
    checker can be handed bit-identical poses."""
    import xreg_b200
    from xreg_b200 import regi

    ctx = xreg_b200.Context(0)

    def render_fixed(det: int, pose: np.ndarray) -> np.ndarray:
        rc = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc.set_volume(vol)
        rc.set_camera_model(synth.make_camera(det))
        rc.set_num_projs(1)
        rc.allocate_resources()
        rc.set_xforms_cam_to_itk_phys([np.asarray(pose, np.float32)])
        rc.compute()
        img = synth.add_noise(rc.proj(0), frac=noise)
        rc.close()
        return img

    def make_objective(det: int, fixed: np.ndarray):
        fn = regi.Intensity2D3DObjFn(ctx, vol, [synth.make_camera(det)], [fixed], metric=metric, max_pop=popsize,
                                     patch_radius=synth.patch_radius_for(det))
        if compose_on_host:
            return (lambda X, pre, post: fn(host_poses(X, pre, post))), fn.close
        return (lambda X, pre, post: fn.eval_se3(X, pre, post)), fn.close

    return make_objective, render_fixed, ctx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--metric", default="grad-ncc", help="grad-ncc (config C3) or patch-grad-ncc")
    ap.add_argument("--pop", type=int, default=100)
    ap.add_argument("--gens", type=int, default=40)
    ap.add_argument("--local-evals", type=int, default=400)
    ap.add_argument("--vol", default="512,512,400")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    nx, ny, nz = (int(v) for v in a.vol.split(","))
    vol = synth.make_volume(nx, ny, nz, spacing=(0.8, 0.8, 1.0))
    truth = synth.nominal_pose(vol).astype(np.float64)
    sp = np.asarray(vol.spacing, np.float64)
    centre = np.asarray(vol.origin, np.float64) + 0.5 * (np.array(vol.dims) - 1.0) * sp
    C4, Ci4 = np.eye(4), np.eye(4)
    C4[:3, 3], Ci4[:3, 3] = centre, -centre
    off = np.array([np.deg2rad(4.0), np.deg2rad(-3.0), np.deg2rad(5.0), 8.0, -6.0, 15.0])
    init = C4 @ exp_se3(off) @ Ci4 @ truth
    make_objective, render_fixed, ctx = gpu_factories(vol, a.metric, a.pop)
    res = run_registration(make_objective, render_fixed, centre, truth, init,
                           levels=((192, "cmaes"), (384, "local"), (768, "local")), popsize=a.pop, cma_gens=a.gens,
                           local_evals=a.local_evals, seed=a.seed, log=lambda r: print(json.dumps(r), flush=True))
    print(json.dumps({"config": "C3: 8x/4x/2x detectors (192/384/768), %s, CMA-ES pop %d x %d generations then local "
                                "refinement (population 1)" % (a.metric, a.pop, a.gens),
                      "volume": [nx, ny, nz], "pose_evals": res["pose_evals"], "seconds": res["seconds"],
                      "pose_evals_per_s": res["pose_evals"] / res["seconds"], "init_error": res["init_error"],
                      "final_error": res["final_error"]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
