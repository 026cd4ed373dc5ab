#!/usr/bin/env python
"""Per-call latency of the objective through the public host API (host poses in, host scalars out)
for the small-population regimes of config C3: BOBYQA (population 1) and CMA-ES (population 100) at the
1536/8, /4, /2 detectors with gradient-NCC, plus patch gradient-NCC.  GPU only; prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xreg_b200  # noqa: E402
from xreg_b200 import regi, synth  # noqa: E402


def main():
    only = os.environ.get("LAT_ONLY")  # "det,metric,pop,reps" -> a single case (for ncu launch lists)
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    nominal = synth.nominal_pose(vol)
    ctx = xreg_b200.Context(0)
    cases = ((192, "grad-ncc"), (384, "grad-ncc"), (768, "grad-ncc"), (192, "patch-grad-ncc"), (384, "patch-grad-ncc"))
    if only:
        cases = ((int(only.split(",")[0]), only.split(",")[1]),)
    for det, metric in cases:
        cam = synth.make_camera(det)
        pops = synth.pose_population(vol, nominal, 100, seed=3)
        rc0 = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc0.set_volume(vol)
        rc0.set_camera_model(cam)
        rc0.set_num_projs(1)
        rc0.allocate_resources()
        rc0.set_xforms_cam_to_itk_phys([nominal])
        rc0.compute()
        fixed = synth.add_noise(rc0.proj(0))
        rc0.close()
        fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric=metric, max_pop=100,
                                     patch_radius=synth.patch_radius_for(det), layout=os.environ.get("LAT_LAYOUT", "default"))
        if os.environ.get("LAT_VARIANT"):   # kernel variant bits (drr.cu launch_pax_k), measurement only
            fn.rc.set_layout_order(int(os.environ["LAT_VARIANT"]) << 1)
        for pop_n in ((1, 100) if not only else (int(only.split(",")[2]),)):
            reps = 200 if pop_n == 1 else 20
            if only:
                reps = int(only.split(",")[3])
            for k in range(5):
                fn(pops[k:k + pop_n])
            t0 = time.perf_counter()
            for k in range(reps):
                fn(pops[(k % 50):(k % 50) + pop_n] if pop_n == 1 else pops)
            dt = (time.perf_counter() - t0) / reps
            # the same evaluation straight through the C ABI (xrc_obj_fn with prebuilt arguments): what a C++ caller
            # pays, without the Python mirror's array conversions and bookkeeping
            import ctypes as C

            from xreg_b200 import _lib
            from xreg_b200.geometry import to12

            lib = _lib.load()
            FP = C.POINTER(C.c_float)
            p12 = [np.ascontiguousarray(to12(pops[(k % 50):(k % 50) + pop_n] if pop_n == 1 else pops)) for k in range(50 if pop_n == 1 else 1)]
            out = np.empty(pop_n, dtype=np.float32)
            outp = out.ctypes.data_as(FP)
            args = [x.ctypes.data_as(FP) for x in p12]
            fn(pops[:pop_n])   # sizes / binds the objects for this population
            for k in range(5):
                _lib.check(lib.xrc_obj_fn(fn.rc.handle, 0, fn._sm_arr, 1, pop_n, args[k % len(args)], outp, None))
            t0 = time.perf_counter()
            for k in range(reps):
                lib.xrc_obj_fn(fn.rc.handle, 0, fn._sm_arr, 1, pop_n, args[k % len(args)], outp, None)
            dt_abi = (time.perf_counter() - t0) / reps
            print(json.dumps({"det": det, "metric": metric, "pop": pop_n, "layout": os.environ.get("LAT_LAYOUT", "default"),
                              "deep": os.environ.get("XRC_PAX_DEEP", ""), "ms_per_call": dt * 1e3,
                              "ms_per_call_c_abi": dt_abi * 1e3, "poses_per_s": pop_n / dt}), flush=True)
        del fn


if __name__ == "__main__":
    main()
