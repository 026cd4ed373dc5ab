#!/usr/bin/env python
"""Probe: does the DRR kernel's time depend on the SIGN of the in-plane / tilt rotation relative to the row / plane pitch
of the PAX stacks (mod 8 records)?  A bank-conflict model of split quarter-warps predicts that it does."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import exp_se3, to12  # noqa: E402

vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
cam = synth.make_camera(480)
nominal = synth.nominal_pose(vol).astype(np.float64)
sp = np.asarray(vol.spacing)
centre = np.asarray(vol.origin) + 0.5 * (np.array(vol.dims) - 1.0) * sp
C4, Ci4 = np.eye(4), np.eye(4)
C4[:3, 3], Ci4[:3, 3] = centre, -centre


def pops(axis, deg, n=100, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        x = rng.standard_normal(6) * [0.002, 0.002, 0.002, 3, 3, 6]
        x[axis] += np.deg2rad(deg)
        out.append((C4 @ exp_se3(x) @ Ci4 @ nominal).astype(np.float32))
    return np.stack(out)


stream = torch.cuda.Stream(device=torch.device("cuda", 0))
cases = [("inplane", 1, 5.0), ("inplane", 1, -5.0), ("tilt_x", 0, 5.0), ("tilt_x", 0, -5.0), ("tilt_z", 2, 5.0), ("tilt_z", 2, -5.0),
         ("none", 1, 0.0)]
with torch.cuda.stream(stream):
    ctx = xreg_b200.Context(0, stream=stream.cuda_stream)
    for pitch in os.environ.get("PITCHES", "default;7,7;0,0;4,4").split(";"):
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
        for name, axis, deg in cases:
            ms = 0.0
            for k in range(4):
                rc.set_poses_array(to12(pops(axis, deg, seed=k)))
                if k == 0:
                    rc.compute()
                    torch.cuda.synchronize()
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                rc.compute()
                e1.record(stream)
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1) / 3
            out["%s%+g" % (name, deg)] = round(ms, 3)
        print(json.dumps(out), flush=True)
        rc.close()
