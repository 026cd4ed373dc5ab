#!/usr/bin/env python
"""Pose-batch sweep of a BASELINE configuration through the public host API, sharded over the ranks of a torchrun job
(one rank per GPU; N = 1 without torchrun): BASELINE config 5, "throughput sweep of pose batch 1-2048 at 1/2/4/8 GPUs"
(768^3 CT, 1536^2 detector, 0.5-voxel step, gradient-NCC), or any other bench.py workload (--workload).

Per batch size: host poses in, host scalars out on every rank (regi.ShardedDeviceObjFn: the (view, pose) list cut into
contiguous balanced chunks, NCCL all-gather of the scalars, one synchronise), wall clock over `reps` calls bracketed by
barriers, max over ranks.  One JSON line per batch size (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048")
    args = ap.parse_args()
    import torch

    import xreg_b200
    from xreg_b200 import regi, synth

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    batches = [int(b) for b in args.batches.split(",")]
    w = dict(bench.WORKLOADS[args.workload])
    w["pop"] = max(batches)
    vol, cams, nominal, pops, held_out = bench.build_scene(w, 2)
    n_views = len(cams)
    step_mm = w.get("step", 1.0)
    ctx = xreg_b200.Context(local)
    rc0 = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc0.set_volume(vol)
    rc0.set_camera_models(cams)
    rc0.set_ray_step_size(step_mm)
    rc0.set_num_projs(n_views)
    rc0.allocate_resources()
    rc0.distribute_xforms_among_cam_models([held_out])
    rc0.compute()
    fixed = [synth.add_noise(rc0.proj(v)) for v in range(n_views)]
    rc0.close()
    # each rank only ever holds its share of the largest batch
    per_rank = (max(batches) * n_views + world - 1) // world
    fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=w["metric"], max_pop=min(max(batches), per_rank),
                                 patch_radius=bench.metric_opts(w), step_size=step_mm)
    sharded = regi.ShardedDeviceObjFn(fn, rank, world)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    ref_vals = None
    for b in batches:
        p = [pops[0][:b], pops[1][:b]]
        reps = 30 if b <= 8 else (6 if b <= 128 else 2)
        for k in range(2):
            out = sharded(p[k % 2])
        barrier()
        t0 = time.perf_counter()
        for k in range(reps):
            out = sharded(p[k % 2])
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert out.shape == (b,) and np.all(np.isfinite(out))
        if rank == 0:
            print(json.dumps({"workload": args.workload, "n_gpus": world, "batch": b, "views": n_views,
                              "ms_per_batch": float(dt.item()) * 1e3, "poses_per_s": b / float(dt.item()),
                              "shares": [hi - lo for lo, hi in regi.unit_chunks(b * n_views, world)][:8],
                              "volume_bytes_resident": fn.rc.volume_bytes(), "first_sim": float(out[0])}), flush=True)
    del sharded
    fn.close()
    ctx.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
