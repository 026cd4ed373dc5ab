#!/usr/bin/env python
"""Benchmark of the DRR + similarity-metric pose-evaluation hot path (one line of JSON per run).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one objective evaluation (Intensity2D3DRegi::obj_fn, xregIntensity2D3DRegi.cpp:571-696) of one pose
population: every view's DRRs of every pose + their metric values + the view mean.  Metric: poses/sec, whole job.

N = 1: the population on one GPU.  N > 1: the SAME population (BASELINE's CMA-ES population of 100 for C2) sharded over
the ranks -- the camera-major (view, pose) projection list cut into N contiguous balanced chunks (100 poses on 8 ranks:
13 13 13 13 12 12 12 12), the CT volume and fixed images replicated per GPU, the per-view scalars all-gathered with NCCL
-- so the 1 -> N curve is STRONG scaling of the reference's population ("scaling": "strong").  The weak-scaling figure
(one whole population per GPU per step) is reported beside it under "weak".
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

FINE = (5.0, 5.0, 5.0, 5.0, 5.0, 10.0)           # fine-stage CMA-ES sigmas (deg, mm), SURVEY 8(d)
COARSE = (15.0, 15.0, 30.0, 50.0, 50.0, 100.0)   # the reference app's first-level sigmas (pelvis...main.cpp:292)

WORKLOADS = {
    "c2": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=480, pop=100, metric="patch-grad-ncc",
               desc="C2: 512x512x400 CT (0.8x0.8x1.0 mm), 480x480 detector, patch gradient-NCC "
                    "(radius 13, Gaussian 5, stride 1), CMA-ES population 100, step 1 mm"),
    "c2-coarse": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=480, pop=100, metric="patch-grad-ncc",
                      sigma=COARSE, desc="C2 with the reference app's coarse CMA-ES sigmas (15,15,30 deg / 50,50,100 mm)"),
    "c2-oblique": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=480, pop=100, metric="patch-grad-ncc",
                       view_rot=35.0, desc="C2 seen 35 degrees off the AP axis (C-arm rotated about the volume's long axis)"),
    "c1": dict(dims=(256, 256, 256), spacing=(1.0, 1.0, 1.0), det=256, pop=1, metric="ncc",
               desc="C1: 256^3 CT, 256x256 detector, NCC, 1 pose"),
    "c3-192": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=192, pop=100, metric="grad-ncc",
                   desc="C3 coarse level: C2 volume, 192x192 detector (8x down-sampled), gradient-NCC, CMA-ES population 100"),
    "c3-768-pop1": dict(dims=(512, 512, 400), spacing=(0.8, 0.8, 1.0), det=768, pop=1, metric="grad-ncc",
                        desc="C3 fine level: C2 volume, 768x768 detector (2x down-sampled), gradient-NCC, BOBYQA population 1"),
    "c4": dict(dims=(512, 512, 512), spacing=(1.0, 1.0, 1.0), det=768, pop=100, metric="patch-grad-ncc",
               views=(0.0, 35.0, -35.0),
               desc="C4: three views (0, +35, -35 deg), 512^3 CT, 768x768 detectors, patch gradient-NCC (radius 21), "
                    "population 100 = 300 DRRs, (view, pose) list sharded across GPUs"),
    "c5": dict(dims=(768, 768, 768), spacing=(1.0, 1.0, 1.0), det=1536, pop=64, metric="grad-ncc", step=0.5,
               desc="C5: 768^3 CT, 1536x1536 detector, 0.5-voxel step, gradient-NCC, pose batch B (--batch, 1..2048)"),
    "small": dict(dims=(96, 96, 80), spacing=(1.0, 1.0, 1.2), det=96, pop=16, metric="patch-grad-ncc",
                  desc="debug: 96x96x80 CT, 96x96 detector, patch gradient-NCC, population 16"),
}
METRIC_NAME = "poses/sec (DRR+patch-GNCC)"


def build_scene(w, n_sets, seed0=0):
    from xreg_b200 import synth

    nx, ny, nz = w["dims"]
    vol = synth.make_volume(nx, ny, nz, spacing=w["spacing"])
    angles = w.get("views", (0.0,))
    src_to_iso = 650.0 if nx >= 256 else 650.0 * 0.55
    if len(angles) > 1:
        cams = synth.multi_view_cameras(w["det"], angles, src_to_iso=src_to_iso)
    else:
        cams = [synth.make_camera(w["det"])]
    if nx < 256:  # debug workload: shrink the geometry so the phantom fills the detector
        from xreg_b200.geometry import CameraModel

        cams = [CameraModel().setup(560.0, w["det"], w["det"], 1.7, 1.7)]
    nominal = synth.nominal_pose(vol, src_to_iso=src_to_iso, view_rot_deg=w.get("view_rot", 0.0))
    sigma = w.get("sigma", FINE)
    pops = [synth.pose_population(vol, nominal, w["pop"], seed=synth.SEED + 17 * (seed0 + k), sigma=sigma)
            for k in range(n_sets)]
    held_out = synth.pose_population(vol, nominal, 1, seed=synth.SEED - 5, sigma=(1, 1, 1, 1, 1, 2))[0]
    return vol, cams, nominal, pops, held_out


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


def metric_opts(w):
    from xreg_b200 import synth

    return synth.patch_radius_for(w["det"])


def cpu_oracle_leg(w, vol, cams, pops, fixed, budget_s, steps=1, warmup=0):
    """Times the CPU restatement (oracle/, OpenMP on all host cores) on a bounded sample of the
    same workload: n poses -> every view's DRRs + metric + view mean.
    Returns (poses_per_sec, cores, n_sample, ms_per_step)."""
    from oracle import xreg_oracle as xo
    from xreg_b200.geometry import to12

    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    xo.set_num_threads(xo.host_cores())
    xcams = [xo.cam_struct(c) for c in cams]
    opts = xo.patch_opts(radius=metric_opts(w))
    step = w.get("step", 1.0)

    def run(poses):
        n = poses.shape[0]
        p12 = np.tile(to12(poses), (len(cams), 1))                     # camera-major (xregRayCastInterface.cpp:97-114)
        ci = np.repeat(np.arange(len(cams), dtype=np.uint32), n)
        d = xo.drr(vol.data, vol.idx_to_phys(), xcams, p12, cam_idx=ci, step_size=step)
        per_view = []
        for v in range(len(cams)):
            dv = d[v * n:(v + 1) * n]
            if w["metric"] == "patch-grad-ncc":
                per_view.append(xo.patch_grad_ncc(fixed[v], dv, opts))
            elif w["metric"] == "grad-ncc":
                per_view.append(xo.grad_ncc(fixed[v], dv))
            else:
                per_view.append(xo.ncc(fixed[v], dv))
        return xo.combine_mean(np.stack(per_view))

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
    vol, cams, nominal, pops, held_out = build_scene(w, max(2, min(args.steps + args.warmup, 8)))
    from oracle import xreg_oracle as xo
    from xreg_b200 import synth
    from xreg_b200.geometry import to12

    xo.set_num_threads(xo.host_cores())   # torchrun exports OMP_NUM_THREADS=1 to every rank
    step = w.get("step", 1.0)
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(c)], to12(held_out[None]), step_size=step)[0])
             for c in cams]
    pps, cores, n, ms = cpu_oracle_leg(w, vol, cams, pops, fixed, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    sample = "%d of %d poses per step (full %dx%d detector x %d view(s), full volume), %d steps" % (
        n, w["pop"], w["det"], w["det"], len(cams), args.steps)
    out = {
        "impl": "reference", "metric": METRIC_NAME, "value": pps, "unit": "poses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "global_batch": w["pop"], "poses_per_step": n,
                   "note": "CPU oracle port of RayCasterLineIntCPU + ImgSimMetric2D*CPU (-O3, no -march=native), OpenMP "
                           "in place of TBB, %d host threads; throughput is linear in poses" % cores},
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
    ap.add_argument("--batch", type=int, default=0, help="override the workload's population / pose batch")
    ap.add_argument("--layout", default="default")
    ap.add_argument("--order", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling figure at N > 1")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "tiles-nccl", "poses"],
                    help="N > 1: shard the detector tiles (projections stored into their owners over NVLink; barrier and gather "
                         "of the scalars by the library's kernel over the peer mappings, or by NCCL) or the poses")
    ap.add_argument("--balance", type=int, default=3,
                    help="N > 1, tile sharding: rounds of clock feedback on the tile plan before the timed region (0 = samples only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])
    if args.batch > 0:
        w["pop"] = args.batch
        w["desc"] += " [batch %d]" % args.batch

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
    # the SAME populations on every rank (strong scaling: the ranks share one population per step)
    vol, cams, nominal, pops, held_out = build_scene(w, n_sets)
    pop_n, n_views = w["pop"], len(cams)
    n_units = pop_n * n_views
    radius = metric_opts(w)
    step_mm = w.get("step", 1.0)
    bounds = regi.unit_chunks(n_units, world)
    u0, u1 = bounds[rank]
    width = max(hi - lo for lo, hi in bounds)
    segs = regi.view_segments(u0, u1, n_views, pop_n)      # (view, first pose, count) runs of this rank's chunk
    n_local = u1 - u0

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = xreg_b200.Context(local_rank, stream=stream.cuda_stream)
        # fixed images: DRRs at a held-out pose + 1% noise, rendered through the public API
        rc0 = xreg_b200.RayCasterLineIntCUDA(ctx, layout=args.layout)
        rc0.set_volume(vol)
        rc0.set_camera_models(cams)
        rc0.set_ray_step_size(step_mm)
        rc0.set_num_projs(n_views)
        rc0.allocate_resources()
        rc0.distribute_xforms_among_cam_models([held_out])
        rc0.compute()
        fixed = [synth.add_noise(rc0.proj(v)) for v in range(n_views)]
        rc0.close()

        # capacity: a whole population per GPU (N = 1 and the weak figure); the strong-scaling steps use width of it
        fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=w["metric"], max_pop=pop_n, patch_radius=radius,
                                     layout=args.layout, step_size=step_mm)
        fn.rc.set_layout_order(args.order)
        npix = cams[0].num_det_rows * cams[0].num_det_cols
        lib = xreg_b200._lib.load()
        check = xreg_b200._lib.check
        import ctypes as C

        # ---- this rank's chunk of every population, resident in HBM (camera-major within the chunk)
        def chunk_arrays(pop, a, b):
            p12 = to12(pop)
            if b <= a:
                return np.zeros((0, 12), np.float32), np.zeros(0, np.uint32)
            rows, cam = [], []
            for v, p0, cnt in regi.view_segments(a, b, n_views, pop_n):
                rows.append(p12[p0:p0 + cnt])
                cam.append(np.full(cnt, v, dtype=np.uint32))
            return np.concatenate(rows), np.concatenate(cam)

        def setup_chunk(a, b):
            """size the ray caster / metrics for units [a, b) and bind each view's metric to its run of projections"""
            fn.rc.set_num_projs(b - a)
            off, active = 0, []
            for v, p0, cnt in regi.view_segments(a, b, n_views, pop_n):
                fn.sims[v].set_num_moving_images(cnt)
                fn.sims[v].set_mov_imgs_buf_from_ray_caster(fn.rc, off)
                active.append(v)
                off += cnt
            fn._cur_pop = -1
            return (C.c_void_p * len(active))(*[fn.sims[v].handle for v in active]), len(active)

        def resident(a, b):
            arr = [chunk_arrays(p, a, b) for p in pops]
            poses_host = np.ascontiguousarray(np.stack([x[0] for x in arr]))
            poses_dev = torch.from_numpy(poses_host).to(dev)
            cam_dev = torch.from_numpy(np.ascontiguousarray(arr[0][1].astype(np.int32))).to(dev)
            return poses_dev, cam_dev, poses_host, arr[0][1]

        def set_resident(r, k, n):
            # device-resident poses; the host mirror only tells the library which volume stacks these poses need
            fn.rc.set_poses_device(r[0][k].data_ptr(), n, r[1].data_ptr(), host_mirror=r[2][k], host_cam_idx=r[3])

        sims_dev = [regi.device_vector(sm.device_sims(), pop_n, dev) for sm in fn.sims]
        send_buf = torch.zeros(max(width, 1), dtype=torch.float32, device=dev)
        gathered = torch.zeros(world * max(width, 1), dtype=torch.float32, device=dev)

        def make_step(a, b, gather):
            res = resident(a, b)
            sm_arr, n_active = setup_chunk(a, b)
            sg = regi.view_segments(a, b, n_views, pop_n)
            wd = max(hi - lo for lo, hi in regi.unit_chunks(n_units, world)) if gather else b - a

            def step(k):
                if b > a:   # a rank without units (fewer units than ranks) only takes part in the gather
                    set_resident(res, k, b - a)
                    check(lib.xrc_eval_batch_async(fn.rc.handle, 0, sm_arr, n_active))
                if gather and world > 1:
                    if len(sg) == 0:
                        send = send_buf[:wd]
                    elif len(sg) == 1:
                        send = sims_dev[sg[0][0]][:wd]      # zero copy: the metric's own result vector
                    else:
                        send, off = send_buf[:wd], 0
                        for v, _, cnt in sg:
                            send[off:off + cnt].copy_(sims_dev[v][:cnt])
                            off += cnt
                    dist.all_gather_into_tensor(gathered[: wd * world], send)
            return step, res

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

        tiles = world > 1 and args.shard in ("tiles", "tiles-nccl")
        sharded, shard_note = None, None
        if world > 1:
            # tile sharding maps every rank's projection buffer into every other rank (CUDA IPC): if a box does not allow
            # that (all ranks must agree), fall back to sharding the poses -- same values, the NCCL path
            ok = torch.ones(1, dtype=torch.float32, device=dev)
            try:
                sharded = regi.ShardedDeviceObjFn(fn, rank, world, mode=args.shard)
                if os.environ.get("XRC_BENCH_FORCE_IPC_FAIL") and tiles and rank == world - 1:
                    raise RuntimeError("forced (XRC_BENCH_FORCE_IPC_FAIL)")     # exercises the fall-back on one rank
            except Exception as exc:   # noqa: BLE001
                ok.zero_()
                shard_note = "tile sharding unavailable on this box (%s): fell back to sharding the poses" % str(exc)[:200]
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                if sharded is not None and tiles:
                    fn.rc.peer_detach()
                if shard_note is None:
                    shard_note = "tile sharding unavailable on another rank: fell back to sharding the poses"
                args.shard, tiles = "poses", False
                sharded = regi.ShardedDeviceObjFn(fn, rank, world, mode="poses")
        flag = torch.zeros(1, dtype=torch.float32, device=dev)

        def make_tile_step():
            """N > 1, tile sharding: every rank ray casts its tiles of ALL units (stored into their owners' buffers over
            NVLink), barrier on the streams, metrics of the units it owns, all-gather of the scalars"""
            res = resident(0, n_units)
            fn.rc.set_num_projs(n_units)
            fn._cur_pop = -1
            sm_all = (C.c_void_p * n_views)(*[sm.handle for sm in fn.sims])

            def drr_only(k):
                set_resident(res, k, n_units)
                fn.rc.compute_tiles()

            def step_ipc(k):
                drr_only(k)
                check(lib.xrc_rc_peer_barrier(fn.rc.handle))
                check(lib.xrc_obj_fn_tiles_enqueue_gather(fn.rc.handle, sm_all, n_views, pop_n))

            def step(k):
                drr_only(k)
                dist.all_reduce(flag)
                if n_local:
                    check(lib.xrc_obj_fn_units_enqueue_metrics(fn.rc.handle, sm_all, n_views, pop_n, u0, n_local))
                if len(segs) == 1:
                    send = sims_dev[segs[0][0]][:width]
                else:
                    send, off = send_buf[:width], 0
                    for v, _, cnt in segs:
                        send[off:off + cnt].copy_(sims_dev[v][:cnt])
                        off += cnt
                dist.all_gather_into_tensor(gathered[: width * world], send)
            return (step if args.shard == "tiles-nccl" else step_ipc), drr_only, res

        # exact sample counts S_k of this rank's share per population -- untimed
        S, F = [], []
        if tiles:
            step_resident, drr_tiles, keep = make_tile_step()
            # one-time set-up, like the plan itself: let the ranks' measured kernel times correct the tile ranges
            tile_plan = sharded.balance(pops[0], rounds=args.balance) if args.balance > 0 else fn.rc.plan_tiles(n_ranks=world)
            fn.rc.set_num_projs(n_units)
            fn._cur_pop = -1
            for k in range(n_sets):
                set_resident(keep, k, n_units)
                a_k, f_k = fn.rc.tile_samples()
                S.append(a_k)
                F.append(f_k)
        else:
            step_resident, keep = make_step(u0, u1, gather=True)
            for k in range(n_sets):
                if n_local == 0:
                    S.append(0)
                    F.append(0)
                    continue
                set_resident(keep, k, n_local)
                S.append(fn.rc.ray_info(counts_only=True)[2])
                F.append(fn.rc.fetched_samples() if args.layout in ("default", "pax") else S[-1])

        for k in range(W):
            step_resident(k)
        with ClockSampler(local_rank) as clk:
            l0 = xreg_b200.launch_count()
            ms_total = timed(step_resident, range(W, W + K))
            launches = xreg_b200.launch_count() - l0
            barrier()
            for sm in fn.sims:
                assert np.all(np.isfinite(regi.device_vector(sm.device_sims(), 1, dev).cpu().numpy()))

            # e2e: the public host API -- host poses in (H2D from pinned staging), host scalars out, every step;
            # N > 1: the sharded objective (chunk per rank, NCCL all-gather of the scalars, D2H, one synchronise)
            e2e_last = [None]

            def step_e2e(k):
                e2e_last[0] = sharded(pops[k]) if world > 1 else fn(pops[k])   # N = 1: the plain one-call objective

            for k in range(W):
                step_e2e(k)
            ms_e2e = timed(step_e2e, range(W, W + K))
            assert e2e_last[0].shape == (pop_n,) and np.all(np.isfinite(e2e_last[0]))

            # dominant kernel alone: K launches of the DRR kernel on this rank's share, CUDA events on its stream
            if tiles:
                fn.rc.set_num_projs(n_units)
                step_drr = drr_tiles
            else:
                setup_chunk(u0, u1)

                def step_drr(k):
                    if n_local:
                        set_resident(keep, k, n_local)
                        fn.rc.compute()

            for k in range(W):
                step_drr(k)
            ms_drr = timed(step_drr, range(W, W + K))
            # same launches with empty-space trimming off (every algorithmic sample fetched)
            fn.rc.set_skip_empty(False)
            for k in range(W):
                step_drr(k)
            ms_drr_dense = timed(step_drr, range(W, W + K))
            fn.rc.set_skip_empty(True)

            weak = None
            if world > 1 and not args.no_weak:
                # secondary: one whole population per GPU per step (what round 1 reported), all-gather included
                step_weak, keep_w = make_step(0, n_units, gather=False)
                for k in range(W):
                    step_weak(k)
                ms_weak = timed(step_weak, range(W, W + K))
                weak = {"value": world * pop_n * K / (ms_weak * 1e-3), "unit": "poses/s", "ms_per_step": ms_weak / K,
                        "global_batch": world * pop_n, "scaling": "weak"}
        clocks = clk.summary()

        total_poses = pop_n * K                                   # whole job: ONE population per step
        value = total_poses / (ms_total * 1e-3)
        e2e_value = total_poses / (ms_e2e * 1e-3)
        S_timed = float(sum(S[W:W + K]))
        F_timed = float(sum(F[W:W + K]))
        R_out = (float(npix) * n_units / world if tiles else float(npix) * n_local) * K
        alg_bytes = 32.0 * S_timed + 4.0 * R_out                  # B_drr = 32 S + 4 R_out, this rank's chunk
        fetched_bytes = 32.0 * F_timed + 4.0 * R_out
        achieved = alg_bytes / (ms_drr * 1e-3) / 1e9
        hbm_peak, peak_src = measured_peak_hbm()
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "c2" and args.layout == "default" and world == 1 and args.batch == 0:
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("drr_dram_bytes_per_launch")
                traffic_src = "static capture: " + tj.get("source", "ncu --set full of this launch, profiles/")
            except Exception:
                traffic = None
        sm_hz = (clocks["sm_mhz"] if clocks else 1965.0) * 1e6
        l1tex_peak = 148 * 128 * sm_hz / 1e9
        vol_bytes = float(vol.data.nbytes)
        roofline = {
            # the gather is served by L1/L2, the binding unit is the L1 data stage / LSU write-back
            # (128 B/clk/SM); HBM is a floor, reported under hbm_*
            "bound": "l1tex", "kernel": "drr_pax_kernel (line-integral ray casting)", "achieved": achieved,
            "peak": l1tex_peak, "unit": "GB/s", "frac": achieved / l1tex_peak,
            "peak_source": "148 SMs x 128 B/clk x sampled SM clock %.0f MHz (L1TEX data stage; SURVEY 8(d) ceiling ii)" % (sm_hz / 1e6),
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": alg_bytes / K, "samples_per_launch": S_timed / K,
            "kernel_ms": ms_drr / K, "kernel_share_of_step": ms_drr / ms_total,
            "note": "frac = algorithmic bytes (32 B per trilinear sample of the reference loop + 4 B per output pixel) / "
                    "kernel time / L1TEX ceiling.  The kernel does not fetch leading/trailing samples a block map proves "
                    "to be zero (bit-identical sums): fetched_* are the samples / bytes it really gathers and fetched_frac "
                    "relates THOSE to the same ceiling; *_no_trim are the same launches with trimming off (fetched = "
                    "algorithmic).  hbm_*: the same algorithmic bytes against the measured HBM copy peak (exceeds 1: cache "
                    "reuse) and the compulsory-traffic floor",
            "fetched_samples_per_launch": F_timed / K,
            "fetched_GBps": fetched_bytes / (ms_drr * 1e-3) / 1e9,
            "fetched_frac": fetched_bytes / (ms_drr * 1e-3) / 1e9 / l1tex_peak,
            "kernel_ms_no_trim": ms_drr_dense / K,
            "achieved_no_trim": alg_bytes / (ms_drr_dense * 1e-3) / 1e9,
            "frac_no_trim": alg_bytes / (ms_drr_dense * 1e-3) / 1e9 / l1tex_peak,
            "hbm_peak": hbm_peak, "hbm_peak_source": peak_src,
            "hbm_algorithmic_over_peak": achieved / hbm_peak,
            "hbm_measured_GBps": (traffic / (ms_drr / K * 1e-3) / 1e9) if traffic else None,
            "hbm_measured_frac": (traffic / (ms_drr / K * 1e-3) / 1e9 / hbm_peak) if traffic else None,
            # SURVEY 8(d) HBM floor: compulsory bytes of one launch = the volume the beams cross (<= the whole f32
            # volume; the PAX stack of the principal axis stores it as 16-byte XY-quad records, 4x) + the projections
            # written + poses + fixed image; measured DRAM traffic / floor = re-read factor
            "hbm_floor_bytes_f32_volume": vol_bytes + 4.0 * npix * n_local + 48 * n_local + 4 * npix,
            "hbm_floor_bytes_pax_stack": 4 * vol_bytes + 4.0 * npix * n_local + 48 * n_local + 4 * npix,
            "reread_factor_vs_f32_volume": (traffic / (vol_bytes + 4.0 * npix * n_local)) if traffic else None,
            "reread_factor_vs_pax_stack": (traffic / (4 * vol_bytes + 4.0 * npix * n_local)) if traffic else None,
            "hbm_floor_frac_of_kernel_time": (4 * vol_bytes + 4.0 * npix * n_local) / (ms_drr / K * 1e-3) / 1e9 / hbm_peak,
        }

        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            pps, cores, n, ms = cpu_oracle_leg(w, vol, cams, pops, fixed, budget_s=20.0)
            cpu = {"value": pps, "unit": "poses/s", "cores": cores, "kind": "port",
                   "sample": "%d of %d poses, full detector(s) and volume, 1 pass (%.1f s)" % (n, pop_n, ms / 1e3)}

        if rank == 0:
            shares = [hi - lo for lo, hi in bounds]
            out = {
                "metric": METRIC_NAME, "value": value, "unit": "poses/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "global_batch": pop_n, "views": n_views,
                           "projections_per_gpu_per_step": shares,
                           "parallelism": (("detector tiles sharded x%d (contiguous ranges of the row-major tile list, cut by measured work and corrected by the clock): every GPU ray casts its tiles of ALL projections "
                                            "and stores them into their owners' buffers over NVLink (peer stores from the DRR "
                                            "kernel, CUDA IPC); the (view, pose) list is cut into contiguous balanced chunks for "
                                            "the metrics; barrier + all-gather of the scalars (%s); volume and fixed images "
                                            "replicated" % (world, "NCCL" if args.shard == "tiles-nccl" else
                                                            "one kernel per rank over the same peer mappings: system-scope flags + NVLink stores")) if tiles else
                                           ("(view, pose) list sharded x%d (contiguous balanced chunks), volume and fixed "
                                            "images replicated, per-view scalars all-gathered (NCCL)" % world)) if world > 1
                                          else "one GPU",
                           "shard": (args.shard if world > 1 else None), "shard_note": shard_note,
                           "tile_plan": ({"bounds": tile_plan, "clock_feedback_rounds": args.balance,
                                          "drr_ms_per_rank_before_last_cut": getattr(sharded, "last_balance_ms", None)}
                                         if tiles else None),
                           "layout": args.layout, "cta_order": args.order,
                           "volume_bytes_resident": fn.rc.volume_bytes(),
                           "cache": "volume payload larger than L2 (126 MB) and a different pose population every step"},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "poses/s", "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": int((n_units if tiles else n_local) * (48 + 4)), "d2h_bytes_per_step": int(width * world * 4) if world > 1 else int(n_units * 4)},
                "gpu_launches": int(launches), "clocks": clocks,
            }
            if weak:
                out["weak"] = weak
            print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
