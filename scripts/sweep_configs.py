#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations other than the headline one (bench.py measures C2):
C1 (256^3 CT, 256^2 detector, NCC), C4 (three views, 512^3 CT, 768^2 detectors, patch gradient-NCC,
population 100) and C5 (768^3 CT, 1536^2 detector, 0.5-voxel step, pose batch 1..2048), through the public
host API (host poses in, host scalars out).  GPU only; one JSON line per case.
SWEEP_CASES=c1,c2bone,c4,c5 selects a subset; SWEEP_C5_MAX caps the C5 batch (default 2048)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import regi, synth  # noqa: E402
from xreg_b200.geometry import to12  # noqa: E402


def render_fixed(ctx, vol, cam, pose, step=1.0):
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_ray_step_size(step)
    rc.set_num_projs(1)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys([pose])
    rc.compute()
    img = synth.add_noise(rc.proj(0))
    rc.close()
    return img


def timed(fn, pops, reps):
    for p in pops[:2]:
        fn(p)
    t0 = time.perf_counter()
    for k in range(reps):
        fn(pops[k % len(pops)])
    return (time.perf_counter() - t0) / reps


def report(case, fn, pops, dt, n_views=1, **extra):
    n = pops[0].shape[0]
    fn(pops[0])  # leaves this population (distributed over the views) in the ray caster
    S = fn.rc.ray_info(counts_only=True)[2]
    F = fn.rc.fetched_samples()
    out = {"case": case, "batch": n, "views": n_views, "ms_per_batch": dt * 1e3, "poses_per_s": n / dt,
           "samples_per_batch": S, "fetched_samples_per_batch": F, "Gsamples_per_s_algorithmic": S / dt / 1e9,
           "drr_GBps_algorithmic": 32.0 * S / dt / 1e9}
    out.update(extra)
    print(json.dumps(out), flush=True)


def main():
    cases = os.environ.get("SWEEP_CASES", "c1,c2bone,c4,c5").split(",")
    ctx = xreg_b200.Context(0)

    if "c1" in cases:
        vol = synth.make_volume(256, 256, 256)
        cam = synth.make_camera(256)
        nominal = synth.nominal_pose(vol)
        fixed = render_fixed(ctx, vol, cam, nominal)
        for pop_n in (1, 100):
            fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="ncc", max_pop=pop_n)
            pops = [synth.pose_population(vol, nominal, pop_n, seed=10 + k) for k in range(4)]
            dt = timed(fn, pops, 200 if pop_n == 1 else 20)
            report("C1 256^3 CT, 256^2 detector, NCC", fn, pops, dt)
            del fn

    if "c2bone" in cases:
        # what the reference's registration apps really ray cast: the CT with everything outside the bone
        # segmentation zeroed (here: the phantom's 12 "bones" only).  Shows what empty-space trimming buys.
        base = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
        vol = xreg_b200.Volume(np.where(base.data >= 0.045, base.data, 0.0).astype(np.float32), base.spacing,
                               base.origin, base.direction)
        cam = synth.make_camera(480)
        nominal = synth.nominal_pose(vol)
        fixed = render_fixed(ctx, vol, cam, nominal)
        fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="patch-grad-ncc", max_pop=100,
                                     patch_radius=synth.patch_radius_for(480))
        pops = [synth.pose_population(vol, nominal, 100, seed=30 + k) for k in range(3)]
        for skip, label in ((1, "on"), (3, "on + interior gaps (on request)"), (0, "off")):
            fn.rc.set_skip_empty(skip)
            dt = timed(fn, pops, 10)
            report("C2 geometry, bone-masked volume (non-bone voxels zero), patch gradient-NCC, trimming %s" % label, fn, pops, dt)
        del fn

    if "c4" in cases:
        vol = synth.make_volume(512, 512, 512)
        cams = synth.multi_view_cameras(768, (0.0, 35.0, -35.0))
        nominal = synth.nominal_pose(vol)
        fixed = []
        for v, cam in enumerate(cams):
            fixed.append(render_fixed(ctx, vol, cam, nominal))
        fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric="patch-grad-ncc", max_pop=100,
                                     patch_radius=synth.patch_radius_for(768))
        pops = [synth.pose_population(vol, nominal, 100, seed=40 + k) for k in range(3)]
        dt = timed(fn, pops, 6)
        report("C4 three views, 512^3 CT, 768^2 detectors, patch gradient-NCC (radius 21)", fn, pops, dt, n_views=3)
        del fn

    if "c5" in cases:
        vol = synth.make_volume(768, 768, 768)
        cam = synth.make_camera(1536)
        nominal = synth.nominal_pose(vol)
        fixed = render_fixed(ctx, vol, cam, nominal, step=0.5)
        bmax = int(os.environ.get("SWEEP_C5_MAX", "2048"))
        # gradient-NCC keeps only the DRRs resident (no gradient images): 2048 x 1536^2 floats = 19.3 GB
        fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="grad-ncc", max_pop=bmax, step_size=0.5)
        for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
            if b > bmax:
                break
            pops = [synth.pose_population(vol, nominal, b, seed=50 + k) for k in range(2)]
            reps = 20 if b <= 8 else (4 if b <= 128 else 1)
            dt = timed(fn, pops, reps)
            report("C5 768^3 CT, 1536^2 detector, 0.5-voxel step, gradient-NCC", fn, pops, dt)
        del fn


if __name__ == "__main__":
    main()
