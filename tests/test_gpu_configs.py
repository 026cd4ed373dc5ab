"""BASELINE.json's configurations C1, C3, C4 and C5 at their full sizes on the GPU: parity with the
oracle on a bounded sample of each (what the oracle finishes in seconds) plus size-independent
properties (batch-order invariance, idempotence, linearity, view-mean).  C2 is in test_gpu_pipeline.py."""
import numpy as np
import pytest

import xreg_b200
from xreg_b200 import regi, synth
from xreg_b200.geometry import downsample_camera_model, to12

pytestmark = pytest.mark.gpu
f32 = np.float32
DRR_REL_TOL = 1.0e-4
SIM_TOL = 1.0e-5


def _drr_check(got, ref, mask):
    np.testing.assert_array_equal(got[mask == 0], ref[mask == 0])
    sel = (mask == 1) & (ref > 0)
    assert sel.any()
    assert np.max(np.abs(got[sel] - ref[sel]) / ref[sel]) <= DRR_REL_TOL


def test_c1_single_drr_256_cubed_ncc(ctx, xo):
    """C1: single line-integral DRR of a 256^3 CT at a 256x256 detector + NCC (the reference's CPU-runnable case)."""
    vol = synth.make_volume(256, 256, 256)
    cam = synth.make_camera(256)
    nominal = synth.nominal_pose(vol)
    pose = synth.pose_population(vol, nominal, 1, sigma=(3, 3, 3, 4, 4, 8))
    xcam = [xo.cam_struct(cam)]
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pose), want_info=True)
    fixed = synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(nominal[None]))[0])
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="ncc", max_pop=1)
    sim = fn(pose)
    _drr_check(fn.rc.proj(0), ref[0], mask[0])
    gmask, gsteps, gS = fn.rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    assert gS == S
    assert abs(sim[0] - xo.ncc(fixed, ref)[0]) <= SIM_TOL
    # the max-intensity kernel of RayCastLineIntParamInterface (xregRayCastInterface.h:575-591) on the same pose
    from xreg_b200 import _lib

    fn.rc.set_kernel_id(_lib.KERNEL_MAX)
    fn.rc.compute()
    mref = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pose), kernel_id=1)
    got = fn.rc.proj(0)
    sel = mask[0] == 1
    assert np.max(np.abs(got[sel] - mref[0][sel])) <= 1e-6 * mref.max()


@pytest.fixture(scope="module")
def c3_scene():
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    nominal = synth.nominal_pose(vol)
    return vol, nominal


@pytest.mark.parametrize("ds", [8, 4, 2])
def test_c3_multires_grad_ncc_population_then_single(ctx, xo, c3_scene, ds):
    """C3: multi-resolution levels (1536 / 8, 4, 2 detectors), gradient-NCC; a CMA-ES population then
    BOBYQA-style single-pose evaluations re-using the same allocation."""
    vol, nominal = c3_scene
    det = 1536 // ds
    cam = synth.make_camera(det)
    xcam = [xo.cam_struct(cam)]
    n_oracle = 3 if det <= 384 else 1
    pop = synth.pose_population(vol, nominal, 100, seed=7 + ds)
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop[:n_oracle]), want_info=True)
    fixed = synth.add_noise(ref[0])
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="grad-ncc", max_pop=100, gauss_width=5)
    sims = fn(pop)
    assert int(np.argmin(sims)) == 0
    for k in range(n_oracle):
        _drr_check(fn.rc.proj(k), ref[k], mask[k])
    assert np.max(np.abs(sims[:n_oracle] - xo.grad_ncc(fixed, ref))) <= SIM_TOL
    # population 1 (NLopt / BOBYQA regime; poses travel in the kernel parameters): same values, bitwise, in any order
    for k in (5, 0, 77):
        one = fn(pop[k:k + 1])
        assert one.shape == (1,) and one[0] == sims[k]
    np.testing.assert_array_equal(fn(pop), sims)


def test_c3_downsampled_camera_matches_reference_rule(xo):
    """DownsampleCameraModel (xregPerspectiveXform.cpp:654-688): detector size and pixel pitch per level."""
    full = synth.make_camera(1536)
    for ds in (8, 4, 2):
        cam = downsample_camera_model(full, 1.0 / ds)
        assert cam.num_det_rows == 1536 // ds and cam.num_det_cols == 1536 // ds
        assert abs(cam.det_col_spacing - 0.194 * ds) < 1e-5
        assert abs(cam.focal_len - full.focal_len) < 1e-2


def test_c4_three_views_512_cubed_patch_gncc(ctx, xo):
    """C4: three simultaneous views (0, +35, -35 degrees), 512^3 CT, 768x768 DRRs, patch gradient-NCC,
    population 100 -> 300 DRRs in one launch, mean over views."""
    vol = synth.make_volume(512, 512, 512)
    cams = synth.multi_view_cameras(768, (0.0, 35.0, -35.0))
    xcams = [xo.cam_struct(c) for c in cams]
    nominal = synth.nominal_pose(vol)
    pop = synth.pose_population(vol, nominal, 100, seed=44)
    radius = synth.patch_radius_for(768)
    assert radius == 21
    # oracle sample: population members 0, 37 and 81 in every view (9 DRRs of 768 x 768)
    members = [0, 37, 81]
    poses, idx = xo.distribute_xforms(to12(pop[members]), 3)
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcams, poses, cam_idx=idx, want_info=True)
    n_m = len(members)
    assert all(mask[v * n_m].mean() > 0.5 for v in range(3))
    assert np.abs(ref[0] - ref[n_m]).max() > 0.1 * ref[0].max()  # the views really differ
    fixed = [synth.add_noise(ref[v * n_m], seed=v) for v in range(3)]
    fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric="patch-grad-ncc", max_pop=100, patch_radius=radius)
    sims = fn(pop)
    assert sims.shape == (100,) and int(np.argmin(sims)) == 0
    for v in range(3):  # view-major buffer: view v's DRR of member m is projection v * pop + m
        for j, m in enumerate(members):
            _drr_check(fn.rc.proj(v * 100 + m), ref[v * n_m + j], mask[v * n_m + j])
    per_view = np.stack([xo.patch_grad_ncc(fixed[v], ref[v * n_m:(v + 1) * n_m], xo.patch_opts(radius=radius)) for v in range(3)])
    assert np.max(np.abs(sims[members] - xo.combine_mean(per_view))) <= SIM_TOL
    perm = np.random.default_rng(1).permutation(100)
    np.testing.assert_array_equal(fn(pop[perm]), sims[perm])
    np.testing.assert_array_equal(fn(pop[40:47]), sims[40:47])


def test_c5_large_volume_full_res_detector_half_voxel_step(ctx, xo):
    """C5: 768^3 CT, 1536x1536 detector, 0.5-voxel step; batch sizes 1..8 give identical per-pose results."""
    vol = synth.make_volume(768, 768, 768)
    cam = synth.make_camera(1536)
    xcam = [xo.cam_struct(cam)]
    nominal = synth.nominal_pose(vol)
    pop = synth.pose_population(vol, nominal, 8, seed=5)
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop[:3]), step_size=0.5, want_info=True)
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_ray_step_size(0.5)
    rc.set_num_projs(8)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(pop))
    rc.compute()
    batch = rc.raw_host_pixel_buf().copy()
    for k in range(3):
        _drr_check(batch[k], ref[k], mask[k])
    gmask, gsteps, gS = rc.ray_info()
    np.testing.assert_array_equal(gmask[:3], mask)
    np.testing.assert_array_equal(gsteps[:3], steps)
    assert int(gsteps[:3].sum()) == S
    assert steps.max() > 1400  # ~768 / 0.5 samples along the central rays
    for n in (1, 3):
        rc.set_num_projs(n)
        rc.set_xforms_cam_to_itk_phys(list(pop[4:4 + n]))
        rc.compute()
        np.testing.assert_array_equal(rc.raw_host_pixel_buf()[:n], batch[4:4 + n])
    rc.close()
