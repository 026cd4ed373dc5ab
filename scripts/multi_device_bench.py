#!/usr/bin/env python
"""Single-process multi-GPU objective (xrc_obj_fn_multi, what a single-threaded C++ caller uses): C2 scene,
population 100 split over 1..N visible devices (strong scaling) and 100 poses per device (weak scaling).
Host API, wall clock around the blocking call; one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import xreg_b200  # noqa: E402
from xreg_b200 import regi, synth  # noqa: E402


def render(ctx, vol, cam, pose):
    rc0 = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc0.set_volume(vol)
    rc0.set_camera_model(cam)
    rc0.set_num_projs(1)
    rc0.allocate_resources()
    rc0.set_xforms_cam_to_itk_phys([pose])
    rc0.compute()
    img = synth.add_noise(rc0.proj(0))
    rc0.close()
    return img


def main_c4():
    """Config C4 (three views, 512^3 CT, 768^2 detectors, patch gradient-NCC r = 21): the (view, pose) list sharded over
    1 / 3 / 4 / ... devices, for the CMA-ES population (100 poses = 300 projections) and the population-1 regime
    (3 projections: one view per device)."""
    n_gpus = torch.cuda.device_count()
    vol = synth.make_volume(512, 512, 512)
    cams = synth.multi_view_cameras(768, (0.0, 35.0, -35.0))
    nominal = synth.nominal_pose(vol)
    ctx = xreg_b200.Context(0)
    fixed = [render(ctx, vol, c, nominal) for c in cams]
    ref = {}
    for n in [d for d in (1, 2, 3, 4, 8) if d <= n_gpus]:
        fn = regi.MultiDeviceObjFn(list(range(n)), vol, cams, fixed, max_pop=100, metric="patch-grad-ncc",
                                   patch_radius=synth.patch_radius_for(768))
        pops = [synth.pose_population(vol, nominal, 100, seed=70 + k) for k in range(3)]
        for pop_n, reps in ((100, 5), (1, 100)):
            for p in pops[:2]:
                fn(p[:pop_n])
            got = fn(pops[0][:pop_n]).copy()
            ref.setdefault(pop_n, got)
            t0 = time.perf_counter()
            for k in range(reps):
                fn(pops[k % 3][:pop_n])
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"config": "C4", "devices": n, "views": 3, "population": pop_n, "ms_per_call": dt * 1e3,
                              "poses_per_s": pop_n / dt,
                              "bitwise_equal_to_one_device": bool(np.array_equal(got, ref[pop_n]))}), flush=True)
        fn.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "c4":
        return main_c4()
    n_gpus = torch.cuda.device_count()
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    cam = synth.make_camera(480)
    nominal = synth.nominal_pose(vol)
    ctx = xreg_b200.Context(0)
    rc0 = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc0.set_volume(vol)
    rc0.set_camera_model(cam)
    rc0.set_num_projs(1)
    rc0.allocate_resources()
    rc0.set_xforms_cam_to_itk_phys([nominal])
    rc0.compute()
    fixed = synth.add_noise(rc0.proj(0))
    rc0.close()
    ref = None
    n = 1
    while n <= n_gpus:
        for mode, pop_n in (("strong", 100), ("weak", 100 * n)):
            if n == 1 and mode == "weak":
                continue
            fn = regi.MultiDeviceObjFn(list(range(n)), vol, [cam], [fixed], max_pop=pop_n, metric="patch-grad-ncc",
                                       patch_radius=synth.patch_radius_for(480))
            pops = [synth.pose_population(vol, nominal, pop_n, seed=70 + k) for k in range(3)]
            for p in pops[:2]:
                out = fn(p)
            if mode == "strong":
                if ref is None:
                    ref = fn(pops[0]).copy()
                same = bool(np.array_equal(fn(pops[0]), ref))
            reps = 10
            t0 = time.perf_counter()
            for k in range(reps):
                fn(pops[k % 3])
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"devices": n, "scaling": mode, "population": pop_n, "ms_per_call": dt * 1e3,
                              "poses_per_s": pop_n / dt, "bitwise_equal_to_one_device": same if mode == "strong" else None}),
                  flush=True)
            fn.close()
        n *= 2


if __name__ == "__main__":
    main()
