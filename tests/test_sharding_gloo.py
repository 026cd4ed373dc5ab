"""World-size-2 CPU test (gloo) of the pose sharding + scalar gather used for multi-GPU runs.
The local evaluation is replaced by the CPU oracle here (test infrastructure standing in for
the per-rank GPU); the product path (Intensity2D3DObjFn) is exercised by the gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from xreg_b200 import regi
from xreg_b200.regi import shard_bounds


def test_shard_bounds():
    assert shard_bounds(100, 8) == [(0, 13), (13, 26), (26, 39), (39, 52), (52, 64), (64, 76), (76, 88), (88, 100)]
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 100, 2048):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_view_segments():
    assert regi.view_segments(0, 300, 3, 100) == [(0, 0, 100), (1, 0, 100), (2, 0, 100)]
    assert regi.view_segments(75, 113, 3, 100) == [(0, 75, 25), (1, 0, 13)]
    assert regi.view_segments(13, 26, 1, 100) == [(0, 13, 13)]
    assert regi.view_segments(5, 5, 2, 4) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_poses, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import xreg_oracle as xo
    from xreg_b200 import synth
    from xreg_b200.geometry import CameraModel, to12
    from xreg_b200.regi import ShardedObjFn

    vol = synth.make_volume(24, 24, 20, spacing=(1.5, 1.5, 1.8))
    cam = CameraModel().setup(300.0, 20, 24, 4.0, 4.0)
    nominal = synth.nominal_pose(vol, src_to_iso=180.0)
    poses = synth.pose_population(vol, nominal, n_poses)
    cams = [xo.cam_struct(cam)]
    fixed = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses[:1]), n_threads=1)[0]

    def local(p):
        if len(p) == 0:
            return np.zeros(0, np.float32)
        return xo.grad_ncc(fixed, xo.drr(vol.data, vol.idx_to_phys(), cams, to12(p), n_threads=1), n_threads=1)

    full = ShardedObjFn(local, rank, world)(poses)
    ref = local(poses)
    q.put((rank, full.tolist(), ref.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_poses", [5, 1])
def test_sharded_objective_world_size_2(n_poses):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_poses, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, ref in res:
        assert len(full) == n_poses
        np.testing.assert_array_equal(np.array(full, np.float32), np.array(ref, np.float32))


def _view_worker(rank, world, port, n_poses, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import xreg_oracle as xo
    from xreg_b200 import synth
    from xreg_b200.geometry import CameraModel, to12
    from xreg_b200.regi import ShardedViewObjFn, combine_mean

    vol = synth.make_volume(24, 24, 20, spacing=(1.5, 1.5, 1.8))
    cams = [xo.cam_struct(CameraModel().setup(f, 20, 24, 4.0, 4.0)) for f in (300.0, 280.0, 320.0)]
    nominal = synth.nominal_pose(vol, src_to_iso=180.0)
    poses = synth.pose_population(vol, nominal, n_poses)
    fixed = [xo.drr(vol.data, vol.idx_to_phys(), [c], to12(poses[:1]), n_threads=1)[0] for c in cams]
    calls = []

    def view_vals(v, p):
        return xo.grad_ncc(fixed[v], xo.drr(vol.data, vol.idx_to_phys(), [cams[v]], to12(p), n_threads=1), n_threads=1)

    def local_units(p, first, count):   # the oracle standing in for Intensity2D3DObjFn.eval_units on this rank
        calls.append((first, count))
        n = len(p)
        out = []
        for v in range(3):
            lo, hi = max(first, v * n), min(first + count, (v + 1) * n)
            if hi > lo:
                out.append(view_vals(v, p[lo - v * n:hi - v * n]))
        return np.concatenate(out)

    fn = ShardedViewObjFn(local_units, 3, rank, world)
    full = fn(poses)
    ref_pv = np.stack([view_vals(v, poses) for v in range(3)])
    q.put((rank, full.tolist(), combine_mean(ref_pv).tolist(), fn.per_view.tolist(), ref_pv.tolist(), calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_poses", [1, 4])
def test_view_sharded_objective_world_size_2(n_poses):
    """Config C4's sharding: the camera-major (view, pose) list split over the ranks, chunks straddling views
    (3 views x 1 pose on 2 ranks -> 2 + 1 projections), scalars all-gathered, view mean on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_view_worker, args=(r, 2, port, n_poses, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = 3 * n_poses
    for rank, full, ref, pv, ref_pv, calls in res:
        assert len(full) == n_poses
        np.testing.assert_array_equal(np.array(full, np.float32), np.array(ref, np.float32))
        np.testing.assert_array_equal(np.array(pv, np.float32), np.array(ref_pv, np.float32))
        first = 0 if rank == 0 else (units + 1) // 2
        count = (units + 1) // 2 if rank == 0 else units // 2
        assert calls == [(first, count)]
