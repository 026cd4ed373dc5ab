"""The sharded objective with the gather on the device (regi.ShardedDeviceObjFn, what bench.py --gpus N times) and the
on-demand PAX stacks.  The 2-rank NCCL run needs two DISTINCT GPUs and is skipped on a one-GPU box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded_device.py -m gpu`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import xreg_b200
from xreg_b200 import regi, synth
from xreg_b200.geometry import CameraModel, to12

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_device_objective_world_1(ctx, xo, small_scene):
    """world_size 1: enqueue + device gather + one synchronise equals the plain objective bit for bit (and the oracle)."""
    vol, cam, nominal = small_scene
    cams = [cam, CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)]
    pop = synth.pose_population(vol, nominal, 9)
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    for cs, fx in ((cams[:1], fixed[:1]), (cams, fixed)):
        single = regi.Intensity2D3DObjFn(ctx, vol, cs, fx, metric="grad-ncc", max_pop=9)
        ref = single(pop)
        single_projs = single.rc.raw_host_pixel_buf().copy()
        fn = regi.Intensity2D3DObjFn(ctx, vol, cs, fx, metric="grad-ncc", max_pop=9)
        for mode in ("poses", "tiles", "tiles-nccl"):     # "tiles" on one rank: this rank's tiles are all tiles, its own buffer the only owner
            sh = regi.ShardedDeviceObjFn(fn, 0, 1, mode=mode)
            np.testing.assert_array_equal(sh(pop), ref)
            np.testing.assert_array_equal(sh(pop[1:4]), ref[1:4])
            np.testing.assert_array_equal(fn.rc.raw_host_pixel_buf()[:3], single_projs[1:4])
            if mode.startswith("tiles"):
                plan = sh.balance(pop, rounds=2, reps=2)     # clock feedback: one rank keeps all tiles, values unchanged
                assert plan[0] == 0 and plan[-1] > 0 and len(plan) == 2
                np.testing.assert_array_equal(sh(pop), ref)
        np.testing.assert_array_equal(sh(pop), ref)
        np.testing.assert_array_equal(sh(pop[2:7]), ref[2:7])
        np.testing.assert_array_equal(sh(pop[4:5]), ref[4:5])
        np.testing.assert_array_equal(fn(pop), ref)          # the plain call still works after the sharded one
        fn.close()
        single.close()


def test_sharded_device_objective_two_gpus_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two distinct GPUs (gpurun --gpus 2)")
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613",
                        os.path.join(ROOT, "tests", "dist", "sharded_device_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"bitwise_equal_to_single_gpu": true' in r.stdout and 'equal_to_single_gpu": false' not in r.stdout
    assert '"mode": "tiles"' in r.stdout and '"mode": "tiles-nccl"' in r.stdout and '"projections_bitwise_equal_to_single_gpu": true' in r.stdout


def test_pax_stacks_are_built_on_demand(ctx, xo):
    """Only the principal-axis stacks the poses can select exist (SURVEY 8(d) HBM budget): one stack for a single view
    (built from the f32 source, which is dropped right after); a second view direction adds its stack, built from the
    first; device-resident poses with a host mirror behave the same, without a mirror all three stacks are built.
    Because every stack a CTA can choose exists before the launch, projections are bitwise the same whichever OTHER
    stacks exist -- and whichever source a stack was built from."""
    import torch

    vol = synth.make_volume(72, 64, 56, spacing=(1.0, 1.1, 1.2))
    cam = CameraModel().setup(420.0, 64, 72, 1.9, 1.9)
    rec = lambda k: 16 * (vol.dims[(k + 1) % 3] + 1) * (vol.dims[(k + 2) % 3] + 1) * (vol.dims[k] + 2)
    f32b = 4 * vol.dims[0] * vol.dims[1] * vol.dims[2]

    def caster():
        rc = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc.set_volume(vol)
        rc.set_camera_model(cam)
        rc.set_num_projs(3)
        rc.allocate_resources()
        return rc

    rc = caster()
    b0 = rc.volume_bytes()
    assert f32b <= b0 < f32b + 4096                      # the source and the (tiny) empty-space map, no stack yet
    ap = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=260.0), 3)
    rc.set_xforms_cam_to_itk_phys(list(ap))
    rc.compute()
    img_ap = rc.raw_host_pixel_buf().copy()
    assert rc.volume_bytes() - (b0 - f32b) == rec(1)     # AP looks along y; the f32 source is gone
    lat = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=90.0), 3)
    rc.set_xforms_cam_to_itk_phys(list(lat))
    rc.compute()
    img_lat = rc.raw_host_pixel_buf().copy()
    assert rc.volume_bytes() - (b0 - f32b) == rec(1) + rec(0)     # lateral looks along x (stack built from stack 1)
    # an oblique population between the two needs both (the 45 degree rays choose per tile)
    obl = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=45.0), 3)
    rc.set_xforms_cam_to_itk_phys(list(obl))
    rc.compute()
    img_obl = rc.raw_host_pixel_buf().copy()
    assert rc.volume_bytes() - (b0 - f32b) == rec(1) + rec(0)

    # a fresh ray caster that sees the oblique population FIRST builds both stacks at once: same bits
    rc2 = caster()
    rc2.set_xforms_cam_to_itk_phys(list(obl))
    rc2.compute()
    np.testing.assert_array_equal(rc2.raw_host_pixel_buf(), img_obl)
    assert rc2.volume_bytes() - (b0 - f32b) == rec(1) + rec(0)
    # device-resident poses with a host mirror: only the stack they need; without: all three, f32 source dropped
    rc3 = caster()
    d_ap = torch.from_numpy(to12(ap)).cuda()
    rc3.set_poses_device(d_ap.data_ptr(), 3, host_mirror=to12(ap))
    rc3.compute()
    np.testing.assert_array_equal(rc3.raw_host_pixel_buf(), img_ap)
    assert rc3.volume_bytes() - (b0 - f32b) == rec(1)
    rc3.set_poses_device(d_ap.data_ptr(), 3)
    rc3.compute()
    np.testing.assert_array_equal(rc3.raw_host_pixel_buf(), img_ap)
    assert rc3.volume_bytes() == b0 - f32b + rec(0) + rec(1) + rec(2)
    d_lat = torch.from_numpy(to12(lat)).cuda()
    rc3.set_poses_device(d_lat.data_ptr(), 3)
    rc3.compute()
    np.testing.assert_array_equal(rc3.raw_host_pixel_buf(), img_lat)
    # oracle parity of what the on-demand stacks produced
    xc = [xo.cam_struct(cam)]
    for imgs, poses in ((img_ap, ap), (img_lat, lat), (img_obl, obl)):
        ref = xo.drr(vol.data, vol.idx_to_phys(), xc, to12(poses))
        sel = ref > 1e-3 * ref.max()
        assert np.max(np.abs(imgs[sel] - ref[sel]) / ref[sel]) <= 1e-4
    for r in (rc, rc2, rc3):
        r.close()
