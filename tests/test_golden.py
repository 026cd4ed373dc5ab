"""Committed golden vectors (tests/golden/small_scene.npz, made by tests/golden/make_golden.py
with the CPU oracle in the build container)."""
import os

import numpy as np
import pytest

import xreg_b200
from tests.golden.make_golden import scene
from xreg_b200 import synth
from xreg_b200.geometry import to12

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_scene.npz"))


def test_oracle_reproduces_golden_vectors(xo):
    vol, cam, poses = scene()
    np.testing.assert_array_equal(poses, G["poses"])
    drr, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    np.testing.assert_array_equal(np.packbits(mask), G["mask"])
    np.testing.assert_array_equal(steps, G["steps"].astype(np.uint32))
    assert S == int(G["S"])
    np.testing.assert_array_equal(drr, G["drr"])
    fixed = G["fixed"]
    o = xo.patch_opts(radius=4)
    np.testing.assert_array_equal(xo.ncc(fixed, drr), G["ncc"])
    np.testing.assert_array_equal(xo.grad_ncc(fixed, drr), G["grad_ncc"])
    np.testing.assert_array_equal(xo.patch_ncc(fixed, drr, o), G["patch_ncc"])
    np.testing.assert_array_equal(xo.patch_grad_ncc(fixed, drr, o), G["patch_grad_ncc"])


@pytest.mark.gpu
def test_cuda_matches_golden_vectors(ctx):
    vol, cam, poses = scene()
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_num_projs(4)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(G["poses"]))
    rc.compute()
    mask, steps, S = rc.ray_info()
    np.testing.assert_array_equal(np.packbits(mask), G["mask"])
    np.testing.assert_array_equal(steps, G["steps"].astype(np.uint32))
    assert S == int(G["S"])
    got, ref = rc.raw_host_pixel_buf(), G["drr"]
    sel = ref != 0
    assert np.all(got[~sel] == 0)
    assert np.max(np.abs(got[sel] - ref[sel]) / ref[sel]) <= 1e-4
    cmask = synth.circular_mask(40, 48)
    cases = [(xreg_b200.ImgSimMetric2DNCCCUDA, None, "ncc"), (xreg_b200.ImgSimMetric2DGradNCCCUDA, None, "grad_ncc"),
             (xreg_b200.ImgSimMetric2DPatchNCCCUDA, None, "patch_ncc"),
             (xreg_b200.ImgSimMetric2DPatchGradNCCCUDA, None, "patch_grad_ncc"),
             (xreg_b200.ImgSimMetric2DPatchGradNCCCUDA, cmask, "patch_grad_ncc_masked"),
             (xreg_b200.ImgSimMetric2DGradNCCCUDA, cmask, "grad_ncc_masked")]
    for cls, m, key in cases:
        sm = cls(ctx)
        sm.set_num_moving_images(4)
        sm.set_fixed_image(G["fixed"])
        sm.set_mov_imgs_buf_from_ray_caster(rc)
        if m is not None:
            sm.set_mask(m)
        if hasattr(sm, "set_patch_radius"):
            sm.set_patch_radius(4)
        sm.allocate_resources()
        sm.compute()
        assert np.max(np.abs(sm.sim_vals() - G[key])) <= 1e-5, key


# ---- round-2 additions: nearest-neighbour interpolation, depth ray caster, projection pre-processing ----------------------
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "round2.npz"))


def test_oracle_reproduces_round2_golden_vectors(xo):
    vol, cam, poses = scene()
    cams = [xo.cam_struct(cam)]
    np.testing.assert_array_equal(xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), interp=1), G2["drr_nn"])
    np.testing.assert_array_equal(xo.depth(vol.data, vol.idx_to_phys(), cams, to12(poses), thresh=float(G2["depth_thresh"]),
                                           n_backtrack=4), G2["depth"])
    np.testing.assert_array_equal(xo.depth(vol.data, vol.idx_to_phys(), cams, to12(poses), interp=1,
                                           thresh=float(G2["depth_nn_thresh"]), n_backtrack=0), G2["depth_nn"])
    assert np.count_nonzero(G2["depth"] < 1e36) > 500
    img, i0 = xo.log_remap(G2["intens"])
    np.testing.assert_array_equal(img, G2["log_remap"])
    assert i0 == G2["log_i0"]
    np.testing.assert_array_equal(xo.downsample_image(G2["intens"], 0.5), G2["down_half"])
    np.testing.assert_array_equal(xo.downsample_image(G2["intens"], 0.25, 0.0), G2["down_quarter_nosmooth"])


@pytest.mark.gpu
def test_cuda_matches_round2_golden_vectors(ctx):
    vol, cam, poses = scene()
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_num_projs(4)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.use_nn_interp()
    rc.compute()
    np.testing.assert_array_equal(rc.raw_host_pixel_buf(), G2["drr_nn"])        # no arithmetic on the voxels: bit-exact
    rc.close()
    dc = xreg_b200.RayCasterDepthCUDA(ctx)
    dc.set_volume(vol)
    dc.set_camera_model(cam)
    dc.set_num_projs(4)
    dc.allocate_resources()
    dc.set_xforms_cam_to_itk_phys(list(poses))
    dc.set_render_thresh(float(G2["depth_thresh"]))
    dc.set_num_backtracking_steps(4)
    dc.compute()
    np.testing.assert_array_equal(dc.raw_host_pixel_buf(), G2["depth"])
    dc.use_nn_interp()
    dc.set_render_thresh(float(G2["depth_nn_thresh"]))
    dc.set_num_backtracking_steps(0)
    dc.compute()
    np.testing.assert_array_equal(dc.raw_host_pixel_buf(), G2["depth_nn"])
    dc.close()
    img, i0 = xreg_b200.log_remap(ctx, G2["intens"])
    assert i0 == G2["log_i0"]
    ulps = np.abs(img.view(np.int32).astype(np.int64) - G2["log_remap"].view(np.int32).astype(np.int64))
    assert ulps.max() <= 1
    np.testing.assert_array_equal(xreg_b200.downsample_image(ctx, G2["intens"], 0.5), G2["down_half"])
    np.testing.assert_array_equal(xreg_b200.downsample_image(ctx, G2["intens"], 0.25, 0.0), G2["down_quarter_nosmooth"])
