#!/usr/bin/env python
"""Probe: DRR kernel time on the C2 workload as a function of the PAX stacks' row / plane pitch modulo one 128-byte
line (XRC_PAX_PITCH="ra,rb": row pitch = ra, plane pitch = rb, both mod 8 records).  GPU only."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import to12  # noqa: E402

vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
cam = synth.make_camera(480)
stream = torch.cuda.Stream(device=torch.device("cuda", 0))
views = {"ap": 0.0, "lateral": 90.0, "oblique": 35.0}
pops = {v: [synth.pose_population(vol, synth.nominal_pose(vol, view_rot_deg=d), 100, seed=100 + k) for k in range(5)]
        for v, d in views.items()}
settings = os.environ.get("PITCHES", "default;0,0;1,1;4,4;4,0;0,4;5,4;3,5;5,3;2,4;4,1;1,4").split(";")
with torch.cuda.stream(stream):
    ctx = xreg_b200.Context(0, stream=stream.cuda_stream)
    for pitch in settings:
        if pitch == "default":
            os.environ.pop("XRC_PAX_PITCH", None)
        else:
            os.environ["XRC_PAX_PITCH"] = pitch
        rc = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc.set_volume(vol)
        rc.set_camera_model(cam)
        rc.set_num_projs(100)
        rc.allocate_resources()
        out = {"pitch": pitch}
        for v in views:
            rc.set_poses_array(to12(pops[v][0]))
            rc.compute()
            torch.cuda.synchronize()
            ms = 0.0
            for k in range(1, 5):
                rc.set_poses_array(to12(pops[v][k]))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                rc.compute()
                e1.record(stream)
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1) / 4
            out[v + "_ms"] = round(ms, 3)
        print(json.dumps(out), flush=True)
        rc.close()
