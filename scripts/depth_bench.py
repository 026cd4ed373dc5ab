#!/usr/bin/env python
"""Timing of the depth ray caster (RayCasterDepthCUDA) on the C2 scene: 512x512x400 CT, 480x480 detector, 100 poses,
bone surface (threshold 0.045 mm^-1), linear / nearest-neighbour interpolation, 0 and 8 refinement steps.  GPU only;
one JSON line per case (wall clock around compute() + a stream synchronise, average of 5 after 2 warm-ups)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import synth  # noqa: E402


def main():
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    cam = synth.make_camera(480)
    nominal = synth.nominal_pose(vol)
    poses = synth.pose_population(vol, nominal, 100, seed=5)
    ctx = xreg_b200.Context(0)
    rc = xreg_b200.RayCasterDepthCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_num_projs(100)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(poses))
    for interp, name in ((0, "linear"), (1, "nearest")):
        for nb in (0, 8):
            rc.set_interp_method(interp)
            rc.set_render_thresh(0.045)
            rc.set_num_backtracking_steps(nb)
            for _ in range(2):
                rc.compute()
            ctx.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                rc.compute()
            ctx.synchronize()
            dt = (time.perf_counter() - t0) / 5
            d = rc.raw_host_pixel_buf()
            hit = d < 1.0e36
            print(json.dumps({"case": "C2 scene, depth of the bone surface (threshold 0.045)", "interp": name, "backtracking_steps": nb,
                              "poses": 100, "ms_per_batch": dt * 1e3, "depth_images_per_s": 100 / dt,
                              "pixels_with_a_surface": float(hit.mean()), "mean_depth_mm": float(d[hit].mean())}), flush=True)
    rc.close()


if __name__ == "__main__":
    main()
