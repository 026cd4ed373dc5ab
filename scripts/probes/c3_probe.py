"""Probe: per-call latency distribution of population-1 objective calls at 384^2 / 768^2, in both orders."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import register_c3 as R
from xreg_b200 import synth

vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
truth = synth.nominal_pose(vol).astype(np.float64)
centre = np.asarray(vol.origin) + 0.5 * (np.array(vol.dims) - 1.0) * np.asarray(vol.spacing)
mk, render, ctx = R.gpu_factories(vol, "grad-ncc", 100)
C4, Ci4 = np.eye(4), np.eye(4); C4[:3, 3], Ci4[:3, 3] = centre, -centre
pre, post = C4.astype(np.float32), (Ci4 @ truth).astype(np.float32)
rng = np.random.default_rng(0)
for det in (384, 768, 384, 192, 384):
    fn, close = mk(det, render(det, truth))
    for mode in ("se3",):
        ts = []
        for k in range(300):
            x = (rng.standard_normal((1, 6)) * [0.01, 0.01, 0.01, 1, 1, 2]).astype(np.float32)
            t0 = time.perf_counter(); fn(x, pre, post); ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(json.dumps({"det": det, "first": ts[0], "min": ts.min(), "median": float(np.median(ts)), "p90": float(np.percentile(ts, 90)),
                          "max": ts.max(), "mean_last100": ts[-100:].mean(), "mean_first100": ts[:100].mean()}), flush=True)
    close()
