#!/usr/bin/env python
"""Benchmark of the DRR + patch gradient-NCC pose-evaluation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one objective evaluation (Intensity2D3DRegi::obj_fn) of a CMA-ES population:
100 poses per GPU -> 100 DRRs (480x480, 512x512x400 CT) + patch gradient-NCC against the
fixed image -> 100 scalars.  Metric: poses/sec, whole job.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, nz, spacing, detector, population, metric)
    "c2": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=480, pop=100, metric="patch-grad-ncc",
               desc="C2: 512x512x400 CT (0.8x0.8x1.0 mm), 480x480 detector, patch gradient-NCC "
                    "(radius 13, Gaussian 5, stride 1), CMA-ES population 100 per GPU, step 1 mm"),
    "c1": dict(dims=(256, 256, 256), spacing=(1.0, 1.0, 1.0), det=256, pop=1, metric="ncc",
               desc="C1: 256^3 CT, 256x256 detector, NCC, 1 pose"),
    "small": dict(dims=(96, 96, 80), spacing=(1.0, 1.0, 1.2), det=96, pop=16, metric="patch-grad-ncc",
                  desc="debug: 96x96x80 CT, 96x96 detector, patch gradient-NCC, population 16"),
}
METRIC_NAME = "poses/sec (DRR+patch-GNCC)"


def build_scene(w, n_sets, seed0=0):
    from xreg_b200 import synth

    nx, ny, nz = w["dims"]
    vol = synth.make_volume(nx, ny, nz, spacing=w["spacing"])
    cam = synth.make_camera(w["det"])
    src_to_iso = 650.0 if nx >= 256 else 650.0 * 0.55
    if nx < 256:  # debug workload: shrink the geometry so the phantom fills the detector
        from xreg_b200.geometry import CameraModel

        cam = CameraModel().setup(560.0, w["det"], w["det"], 1.7, 1.7)
    nominal = synth.nominal_pose(vol, src_to_iso=src_to_iso)
    pops = [synth.pose_population(vol, nominal, w["pop"], seed=synth.SEED + 17 * (seed0 + k)) for k in range(n_sets)]
    held_out = synth.pose_population(vol, nominal, 1, seed=synth.SEED - 5, sigma=(1, 1, 1, 1, 1, 2))[0]
    return vol, cam, nominal, pops, held_out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_oracle_leg(w, vol, cam, pops, fixed, budget_s, steps=1, warmup=0):
    """Times the CPU restatement (oracle/, OpenMP on all host cores) on a bounded sample of the
    same workload: n poses -> DRR + metric.  Returns (poses_per_sec, cores, n_sample, ms_per_step)."""
    from oracle import xreg_oracle as xo
    from xreg_b200 import synth
    from xreg_b200.geometry import to12

    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    xo.set_num_threads(xo.host_cores())
    xcam = [xo.cam_struct(cam)]
    radius = synth.patch_radius_for(w["det"])
    opts = xo.patch_opts(radius=radius)

    def run(poses):
        d = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(poses))
        if w["metric"] == "patch-grad-ncc":
            return xo.patch_grad_ncc(fixed, d, opts)
        return xo.ncc(fixed, d)

    t0 = time.perf_counter()
    run(pops[0][:1])
    t1 = time.perf_counter() - t0
    n = int(max(1, min(w["pop"], budget_s / max(t1, 1e-3) / max(1, steps + warmup))))
    for k in range(warmup):
        run(pops[k % len(pops)][:n])
    t0 = time.perf_counter()
    for k in range(steps):
        run(pops[(warmup + k) % len(pops)][:n])
    dt = time.perf_counter() - t0
    return n * steps / dt, xo.num_threads(), n, 1e3 * dt / steps


def run_reference(args, w):
    """--impl reference: the reference's CPU implementation of the path.  Its own sources cannot be
    compiled here (ITK/Eigen/OpenCV/TBB absent), so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vol, cam, nominal, pops, held_out = build_scene(w, max(2, min(args.steps + args.warmup, 8)))
    from oracle import xreg_oracle as xo
    from xreg_b200 import synth
    from xreg_b200.geometry import to12

    xo.set_num_threads(xo.host_cores())   # torchrun exports OMP_NUM_THREADS=1 to every rank
    fixed = synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(held_out[None]))[0])
    pps, cores, n, ms = cpu_oracle_leg(w, vol, cam, pops, fixed, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    sample = "%d of %d poses per step (full %dx%d detector, full volume), %d steps" % (n, w["pop"], w["det"], w["det"], args.steps)
    out = {
        "impl": "reference", "metric": METRIC_NAME, "value": pps, "unit": "poses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "poses_per_step": n, "note": "CPU oracle port of RayCasterLineIntCPU + "
                   "ImgSimMetric2DPatchGradNCCCPU, OpenMP in place of TBB; throughput is linear in poses"},
        "cpu_baseline": {"value": pps, "unit": "poses/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--layout", default="default")
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch

    import xreg_b200
    from xreg_b200 import regi, synth
    from xreg_b200.geometry import to12

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU leg)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, args.warmup
    n_sets = K + W
    # every rank evaluates its own populations (weak scaling: 100 poses per GPU per step)
    vol, cam, nominal, pops, held_out = build_scene(w, n_sets, seed0=1000 * rank)
    pop_n = w["pop"]
    radius = synth.patch_radius_for(w["det"])

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = xreg_b200.Context(local_rank, stream=stream.cuda_stream)
        # fixed image: DRR at a held-out pose + 1% noise, rendered through the public API
        rc0 = xreg_b200.RayCasterLineIntCUDA(ctx, layout=args.layout)
        rc0.set_volume(vol)
        rc0.set_camera_model(cam)
        rc0.set_num_projs(1)
        rc0.allocate_resources()
        rc0.set_xforms_cam_to_itk_phys([held_out])
        rc0.compute()
        fixed = synth.add_noise(rc0.proj(0))
        rc0.close()

        fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric=w["metric"], max_pop=pop_n, patch_radius=radius,
                                     layout=args.layout)
        fn.rc.set_layout_order(args.order)
        sm = fn.sims[0]
        npix = cam.num_det_rows * cam.num_det_cols

        # exact sample counts S_k per population -- untimed
        S, F = [], []
        for k in range(n_sets):
            fn.rc.set_poses_array(to12(pops[k]))
            S.append(fn.rc.ray_info(counts_only=True)[2])
            F.append(fn.rc.fetched_samples() if args.layout in ("default", "pax") else S[-1])

        poses_host = np.ascontiguousarray(np.stack([to12(p) for p in pops]))          # (n_sets, pop, 12)
        poses_dev = torch.from_numpy(poses_host).to(dev)                              # resident in HBM
        sims_ptr = sm.device_sims()

        class _Ptr:  # zero-copy torch view of the metric's device result vector
            __cuda_array_interface__ = {"shape": (pop_n,), "typestr": "<f4", "data": (sims_ptr, False), "version": 2}

        sims_dev = torch.as_tensor(_Ptr(), device=dev)
        gathered = torch.empty(world * pop_n, dtype=torch.float32, device=dev) if world > 1 else None
        lib = xreg_b200._lib.load()
        sm_arr = (__import__("ctypes").c_void_p * 1)(sm.handle)

        def step_resident(k):
            fn.rc.set_poses_device(poses_dev[k].data_ptr(), pop_n)
            xreg_b200._lib.check(lib.xrc_eval_batch_async(fn.rc.handle, 0, sm_arr, 1))
            if world > 1:
                dist.all_gather_into_tensor(gathered, sims_dev)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        def timed(fn_step, ks):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in ks:
                fn_step(k)
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        for k in range(W):
            step_resident(k)
        with ClockSampler(local_rank) as clk:
            l0 = xreg_b200.launch_count()
            ms_total = timed(step_resident, range(W, W + K))
            launches = xreg_b200.launch_count() - l0
            sims_last = sims_dev.clone().cpu().numpy()
            assert np.all(np.isfinite(sims_last))

            # e2e: public API with HOST buffers -- poses H2D from pinned staging, sims D2H, every step
            sharded = regi.ShardedObjFn(lambda p: fn(p), rank=0, world_size=1)

            def step_e2e(k):
                s = sharded(pops[k])
                if world > 1:
                    dist.all_gather_into_tensor(gathered, torch.from_numpy(s).to(dev))

            for k in range(W):
                step_e2e(k)
            ms_e2e = timed(step_e2e, range(W, W + K))

            # dominant kernel alone: K launches of the DRR kernel, CUDA events on its stream
            for k in range(W):
                fn.rc.set_poses_device(poses_dev[k].data_ptr(), pop_n)
                fn.rc.compute()

            def step_drr(k):
                fn.rc.set_poses_device(poses_dev[k].data_ptr(), pop_n)
                fn.rc.compute()

            ms_drr = timed(step_drr, range(W, W + K))
            # same launches with empty-space trimming off (every algorithmic sample fetched)
            fn.rc.set_skip_empty(False)
            for k in range(W):
                step_drr(k)
            ms_drr_dense = timed(step_drr, range(W, W + K))
            fn.rc.set_skip_empty(True)
        clocks = clk.summary()

        total_poses = world * pop_n * K
        value = total_poses / (ms_total * 1e-3)
        e2e_value = total_poses / (ms_e2e * 1e-3)
        S_timed = float(sum(S[W:W + K]))
        F_timed = float(sum(F[W:W + K]))
        alg_bytes = 32.0 * S_timed + 4.0 * npix * pop_n * K           # B_drr = 32 S + 4 R_out
        achieved = alg_bytes / (ms_drr * 1e-3) / 1e9
        peak, peak_src = measured_peak_hbm()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "c2" and args.layout == "default":   # the capture is of the C2 launch
            try:
                traffic = json.load(open(tpath)).get("drr_dram_bytes_per_launch")
            except Exception:
                traffic = None
        sm_hz = (clocks["sm_mhz"] if clocks else 1965.0) * 1e6
        l1tex_peak = 148 * 128 * sm_hz / 1e9
        roofline = {
            "bound": "hbm", "kernel": "drr_kernel (line-integral ray casting)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes / K, "samples_per_launch": S_timed / K,
            "kernel_ms": ms_drr / K, "kernel_share_of_step": ms_drr / ms_total,
            "note": "algorithmic bytes = 32 B per trilinear sample of the reference loop + 4 B per output pixel; the "
                    "gather is served by L1/L2 (volume >> L2 but beams overlap), so the fraction of the HBM copy peak "
                    "may exceed 1.  The kernel does not fetch leading/trailing samples a block map proves to be zero "
                    "(bit-identical sums): fetched_* are the samples / bytes it really gathers, l1tex_frac relates "
                    "THOSE bytes to the 148 SM x 128 B/clk L1 ceiling at the sampled SM clock, and *_no_trim are the "
                    "same launches with trimming off (fetched = algorithmic)",
            "fetched_samples_per_launch": F_timed / K,
            "fetched_GBps": (32.0 * F_timed + 4.0 * npix * pop_n * K) / (ms_drr * 1e-3) / 1e9,
            "l1tex_peak": l1tex_peak,
            "l1tex_frac": (32.0 * F_timed + 4.0 * npix * pop_n * K) / (ms_drr * 1e-3) / 1e9 / l1tex_peak,
            # SURVEY 8(d) HBM floor: compulsory bytes of one launch = the volume the beams cross (<= the whole f32
            # volume; the PAX stack of the principal axis stores it as 16-byte XY-quad records, 4x) + the projections
            # written + poses + fixed image; measured DRAM traffic / floor = re-read factor
            "hbm_floor_bytes_f32_volume": float(vol.data.nbytes + 4 * npix * pop_n + 48 * pop_n + 4 * npix),
            "hbm_floor_bytes_pax_stack": float(4 * vol.data.nbytes + 4 * npix * pop_n + 48 * pop_n + 4 * npix),
            "reread_factor_vs_f32_volume": (traffic / float(vol.data.nbytes + 4 * npix * pop_n)) if traffic else None,
            "reread_factor_vs_pax_stack": (traffic / float(4 * vol.data.nbytes + 4 * npix * pop_n)) if traffic else None,
            "hbm_floor_frac_of_kernel_time": (4 * vol.data.nbytes + 4 * npix * pop_n) / (ms_drr / K * 1e-3) / 1e9 / peak,
            "kernel_ms_no_trim": ms_drr_dense / K,
            "achieved_no_trim": alg_bytes / (ms_drr_dense * 1e-3) / 1e9,
            "l1tex_frac_no_trim": alg_bytes / (ms_drr_dense * 1e-3) / 1e9 / l1tex_peak,
        }

        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            pps, cores, n, ms = cpu_oracle_leg(w, vol, cam, pops, fixed, budget_s=20.0)
            cpu = {"value": pps, "unit": "poses/s", "cores": cores, "kind": "port",
                   "sample": "%d of %d poses, full detector and volume, 1 pass (%.1f s)" % (n, pop_n, ms / 1e3)}

        if rank == 0:
            out = {
                "metric": METRIC_NAME, "value": value, "unit": "poses/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "global_batch": world * pop_n, "poses_per_gpu_per_step": pop_n,
                           "parallelism": "pose-sharded x%d, volume replicated, scalars all-gathered" % world,
                           "layout": args.layout, "cta_order": args.order,
                           "cache": "volume payload larger than L2 (126 MB) and a different pose population every step"},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "poses/s", "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": int(pop_n * (48 + 4)), "d2h_bytes_per_step": int(pop_n * 4)},
                "gpu_launches": int(launches), "clocks": clocks,
            }
            print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
