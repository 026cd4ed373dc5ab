"""Run under torch.distributed.run (one rank per GPU, NCCL): the (view, pose)-sharded device objective
(regi.ShardedDeviceObjFn: chunk per rank, all-gather of the scalars out of the metrics' device vectors) must give,
on every rank, the single-GPU objective's values BIT FOR BIT -- single view and three views, uneven chunks, fewer
units than ranks.  Prints one JSON line per rank-0 case; exit code 0 = all equal."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import xreg_b200
    from xreg_b200 import regi, synth
    from xreg_b200.geometry import CameraModel

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert torch.cuda.device_count() >= world, "one distinct GPU per rank"

    vol = synth.make_volume(64, 64, 48, spacing=(0.9, 1.1, 1.3))
    cam = CameraModel().setup(400.0, 80, 96, 1.6, 1.5)
    cams3 = [cam, CameraModel().setup(380.0, 80, 96, 1.7, 1.4), CameraModel().setup(420.0, 80, 96, 1.5, 1.6)]
    nominal = synth.nominal_pose(vol, src_to_iso=250.0)
    pop = synth.pose_population(vol, nominal, 11)
    ctx = xreg_b200.Context(local)
    ok = True
    for cams, metric in (([cam], "patch-grad-ncc"), (cams3, "grad-ncc"), ([cam], "ncc")):
        rc0 = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc0.set_volume(vol)
        rc0.set_camera_models(cams)
        rc0.set_num_projs(len(cams))
        rc0.allocate_resources()
        rc0.distribute_xforms_among_cam_models([pop[0]])
        rc0.compute()
        fixed = [synth.add_noise(rc0.proj(v), seed=v) for v in range(len(cams))]
        rc0.close()
        single = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=metric, max_pop=11, patch_radius=6)
        ref = single(pop)
        for mode in ("poses", "tiles", "tiles-nccl"):
            fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=metric, max_pop=11, patch_radius=6)
            sharded = regi.ShardedDeviceObjFn(fn, rank, world, mode=mode)
            for i_sel, sel in enumerate((slice(0, 11), slice(3, 10), slice(5, 6), slice(0, 2), slice(0, 11))):
                if mode.startswith("tiles") and i_sel == 3:
                    # clock feedback on the tile plan (collective): the plan may move, the values may not
                    plan = sharded.balance(pop, rounds=2, reps=2)
                    assert len(plan) == world + 1 and plan[0] == 0 and all(a <= b for a, b in zip(plan, plan[1:])), plan
                got = sharded(pop[sel])
                same = bool(np.array_equal(got, ref[sel]))
                # every rank must hold the same full vector
                t = torch.from_numpy(got.copy()).to(dev)
                lst = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(lst, t)
                same = same and all(bool(torch.equal(lst[0], x)) for x in lst)
                ok = ok and same
                if rank == 0:
                    print(json.dumps({"world": world, "mode": mode, "views": len(cams), "metric": metric, "poses": int(len(got)),
                                      "bitwise_equal_to_single_gpu": same}), flush=True)
            if mode.startswith("tiles"):
                # the projections this rank OWNS were assembled from every rank's tiles: they must equal the single-GPU ones
                n = 11
                b, e = regi.unit_chunks(len(cams) * n, world)[rank]
                mine = fn.rc.raw_host_pixel_buf()[b:e]
                same = bool(np.array_equal(mine, single.rc.raw_host_pixel_buf()[b:e])) if e > b else True
                ok = ok and same
                print(json.dumps({"world": world, "mode": mode, "rank": rank, "owned_projections": [b, e],
                                  "projections_bitwise_equal_to_single_gpu": same}), flush=True)
                dist.barrier()
                fn.rc.peer_detach()
            del sharded
            fn.close()
        single.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    code = 0 if int(flag.item()) == 1 else 1
    ctx.close()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(code)   # skip interpreter teardown: pinned buffers would be released after the CUDA context is gone


if __name__ == "__main__":
    main()
