#!/usr/bin/env python
"""DRR kernel time for every volume layout x CTA order on the C2 workload (GPU only).
Writes gpurun_out/sweep_layouts.json.  Not a bench value: it picks defaults."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import to12  # noqa: E402


def main():
    det = int(os.environ.get("SWEEP_DET", "480"))
    pop_n = int(os.environ.get("SWEEP_POP", "100"))
    sigmas = {"fine": (5, 5, 5, 5, 5, 10), "coarse": (15, 15, 30, 50, 50, 100)}
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    cam = synth.make_camera(det)
    views = {"ap": 0.0, "lateral": 90.0, "oblique": 35.0}
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    results = []
    only = os.environ.get("SWEEP_ONLY")  # e.g. "quad,lateral,fine,0" (for ncu captures)
    layouts = os.environ.get("SWEEP_LAYOUTS", "pax").split(",")
    if only:
        layouts = [only.split(",")[0]]
    with torch.cuda.stream(stream):
        ctx = xreg_b200.Context(0, stream=stream.cuda_stream)
        for layout in layouts:
            rc = xreg_b200.RayCasterLineIntCUDA(ctx, layout=layout)
            rc.set_volume(vol)
            rc.set_camera_model(cam)
            rc.set_num_projs(pop_n)
            rc.allocate_resources()
            for vname, vdeg in views.items():
                nominal = synth.nominal_pose(vol, view_rot_deg=vdeg)
                for sname, sig in sigmas.items():
                    if sname == "coarse" and vname != "ap":
                        continue
                    pops = [synth.pose_population(vol, nominal, pop_n, seed=100 + k, sigma=sig) for k in range(6)]
                    for order in [int(o) for o in os.environ.get("SWEEP_ORDERS", "0").split(",")]:
                        if only and (vname, sname, str(order)) != tuple(only.split(",")[1:4]):
                            continue
                        rc.set_layout_order(order)
                        S = 0
                        for k in range(2):
                            rc.set_poses_array(to12(pops[k]))
                            rc.compute()
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ms = 0.0
                        for k in range(2, 6):
                            rc.set_poses_array(to12(pops[k]))
                            S += rc.ray_info(counts_only=True)[2]
                            torch.cuda.synchronize()
                            e0.record(stream)
                            rc.compute()
                            e1.record(stream)
                            torch.cuda.synchronize()
                            ms += e0.elapsed_time(e1)
                        ms /= 4
                        S /= 4
                        r = dict(layout=layout, view=vname, sigma=sname, order=order, ms=ms, samples=S,
                                 gsamples_per_s=S / ms / 1e6, poses_per_s=pop_n / ms * 1e3)
                        results.append(r)
                        print(json.dumps(r), flush=True)
            rc.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "sweep_layouts.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
