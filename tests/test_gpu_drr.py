"""GPU parity of RayCasterLineIntCUDA against the CPU oracle (through the C ABI):
bit-exact clip masks / step counts / pixel indexing, per-pixel DRR relative error <= 1e-4."""
import numpy as np
import pytest

import xreg_b200
from xreg_b200 import synth
from xreg_b200.geometry import CameraModel, Volume, to12

pytestmark = pytest.mark.gpu
f32 = np.float32

DRR_REL_TOL = 1.0e-4  # BASELINE.json north_star: per-pixel DRR relative error <= 1e-4
LAYOUTS = ["linear", "quad", "oct", "tex", "tex_quad"]  # one lerp order: bitwise equal
ALL_LAYOUTS = LAYOUTS + ["pax"]


def _make_rc(ctx, vol, cams, n, layout="default"):
    rc = xreg_b200.RayCasterLineIntCUDA(ctx, layout=layout)
    rc.set_volume(vol)
    rc.set_camera_models(cams)
    rc.set_num_projs(n)
    rc.allocate_resources()
    return rc


def _check_drr(got, ref, mask):
    """relative error on marched pixels; exact zeros elsewhere"""
    assert np.all(got[mask == 0] == ref[mask == 0])
    sel = (mask == 1) & (np.abs(ref) > 0)
    if sel.any():
        rel = np.abs(got[sel] - ref[sel]) / np.abs(ref[sel])
        assert rel.max() <= DRR_REL_TOL, rel.max()
    return float(np.abs(got - ref).max())


@pytest.mark.parametrize("layout", ALL_LAYOUTS)
def test_drr_parity_all_layouts(ctx, xo, small_scene, layout):
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, 5, sigma=(10, 10, 10, 6, 6, 6))
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    rc = _make_rc(ctx, vol, [cam], 5, layout)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    got = rc.raw_host_pixel_buf()
    gmask, gsteps, gS = rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)      # bit-exact intersection masks
    np.testing.assert_array_equal(gsteps, steps)    # identical sample counts per ray
    assert gS == S
    _check_drr(got, ref, mask)
    np.testing.assert_array_equal(rc.proj(3), got[3])


@pytest.mark.parametrize("layout", ["pax", "linear", "quad", "oct"])
@pytest.mark.parametrize("kernel_id", [0, 1])
def test_nearest_neighbour_interpolation_is_bit_exact(ctx, xo, small_scene, layout, kernel_id):
    """kRAY_CAST_INTERP_NN (xregRayCastLineIntCPU.cpp:128-130): no arithmetic on the voxel values, so the sums (and the
    max kernel's maxima) equal the oracle's bit for bit -- for every payload the voxel can be read from, several views
    (the on-demand stack differs), REPLACE and ACCUM; switching back to linear gives the linear result again."""
    vol, cam, nominal = small_scene
    xc = [xo.cam_struct(cam)]
    rc = _make_rc(ctx, vol, [cam], 4, layout)
    rc.set_kernel_id(kernel_id)
    for view in (0.0, 90.0, 40.0):
        poses = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=250.0, view_rot_deg=view), 4,
                                      sigma=(10, 10, 10, 6, 6, 6))
        ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xc, to12(poses), want_info=True, interp=1, kernel_id=kernel_id)
        lin = xo.drr(vol.data, vol.idx_to_phys(), xc, to12(poses), kernel_id=kernel_id)
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.use_nn_interp()
        rc.compute()
        got = rc.raw_host_pixel_buf().copy()
        assert got.tobytes() == ref.tobytes() and got.tobytes() != lin.tobytes()
        gmask, gsteps, gS = rc.ray_info()
        np.testing.assert_array_equal(gmask, mask)
        assert gS == S
        rc.use_proj_store_accum_method()
        rc.compute()
        ref2 = xo.drr(vol.data, vol.idx_to_phys(), xc, to12(poses), interp=1, kernel_id=kernel_id, buf=ref.copy())
        assert rc.raw_host_pixel_buf().tobytes() == ref2.tobytes()
        rc.use_proj_store_replace_method()
        rc.use_linear_interp()
        rc.compute()
        _check_drr(rc.raw_host_pixel_buf(), lin, mask)
    rc.use_sinc_interp()
    with pytest.raises(xreg_b200.UnsupportedOperationException):
        rc.compute()
    rc.close()


@pytest.mark.parametrize("layout", ["pax", "linear", "quad"])
def test_depth_ray_caster_is_bit_exact(ctx, xo, small_scene, layout):
    """RayCasterDepthCUDA == RayCasterDepthCPU (oracle xo_depth, itself pinned to the reference's RayCastDepthFn), bit for
    bit: thresholds, step-halving refinement, linear and nearest-neighbour interpolation, min on top of the default
    background (kRAY_CAST_MAX_DEPTH), of a previous depth image (ACCUM) and of a background projection; two cameras."""
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    xc = [xo.cam_struct(cam), xo.cam_struct(cam2)]
    vmax = float(vol.data.max())
    rc = xreg_b200.RayCasterDepthCUDA(ctx, layout=layout)
    rc.set_volume(vol)
    rc.set_camera_models([cam, cam2])
    rc.set_num_projs(6)
    rc.allocate_resources()
    cam_idx = np.array([0, 0, 0, 1, 1, 1], np.uint32)
    found = 0
    for view in (0.0, 90.0, 40.0):
        pop = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=250.0, view_rot_deg=view), 3, sigma=(10, 10, 10, 6, 6, 6))
        rc.distribute_xforms_among_cam_models(list(pop))
        p12 = to12(np.concatenate([pop, pop]))
        for interp in (0, 1):
            rc.set_interp_method(interp)
            for frac, nb, step in ((0.5, 0, 1.0), (0.7, 6, 1.0), (0.2, 20, 0.6), (3.0, 2, 1.0)):
                rc.set_render_thresh(frac * vmax)
                rc.set_num_backtracking_steps(nb)
                rc.set_ray_step_size(step)
                rc.use_proj_store_replace_method()
                rc.compute()
                got = rc.raw_host_pixel_buf().copy()
                ref = xo.depth(vol.data, vol.idx_to_phys(), xc, p12, cam_idx=cam_idx, step_size=step, interp=interp,
                               thresh=frac * vmax, n_backtrack=nb)
                assert got.tobytes() == ref.tobytes(), (layout, view, interp, frac, nb)
                found += int(np.count_nonzero(ref < 1.0e36))
                if frac > 1.0:
                    assert np.all(got == np.float32(1.0e37))
                # ACCUM: min with what is there (a second surface at a lower threshold can only come closer)
                rc.use_proj_store_accum_method()
                rc.set_render_thresh(0.5 * frac * vmax)
                rc.compute()
                ref2 = xo.depth(vol.data, vol.idx_to_phys(), xc, p12, cam_idx=cam_idx, step_size=step, interp=interp,
                                thresh=0.5 * frac * vmax, n_backtrack=nb, buf=ref.copy())
                got2 = rc.raw_host_pixel_buf()
                assert got2.tobytes() == ref2.tobytes() and np.all(got2 <= got)
    assert found > 1000
    rc.use_sinc_interp()
    with pytest.raises(xreg_b200.UnsupportedOperationException):
        rc.compute()
    rc.close()


def test_layouts_agree_bitwise(ctx, small_scene):
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, 3)
    outs = []
    for layout in LAYOUTS:
        rc = _make_rc(ctx, vol, [cam], 3, layout)
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.compute()
        outs.append(rc.raw_host_pixel_buf())
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])
    # the principal-axis stacks lerp in a different order (b, c, a): same samples, last-ulp differences only
    rc = _make_rc(ctx, vol, [cam], 3, "pax")
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    pax = rc.raw_host_pixel_buf()
    assert np.array_equal(pax == 0, outs[0] == 0)
    sel = outs[0] != 0
    assert np.max(np.abs(pax[sel] - outs[0][sel]) / np.abs(outs[0][sel])) < 2e-6


@pytest.mark.parametrize("view_deg,axis", [(0.0, "y"), (90.0, "x"), (35.0, "y/x"), (55.0, "x/y")])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])  # bit0: packed f32x2 instead of scalar FP32; bit1: force the clamped loop
def test_pax_every_stack_and_variant(ctx, xo, view_deg, axis, variant):
    """Every principal-axis stack (rays along y, x and oblique; a tilted C-arm for z), packed and
    scalar arithmetic, fast and clamped marching loops: all must match the oracle."""
    vol = synth.make_volume(40, 56, 48, spacing=(1.1, 0.9, 1.0))
    cam = CameraModel().setup(400.0, 48, 40, 2.2, 2.2)
    for tilt in (None, "z"):
        nominal = synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=view_deg)
        if tilt == "z":  # rotate the C-arm about x so that rays run along the volume's z axis
            rot = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=f32)
            c = np.asarray(vol.origin) + 0.5 * (np.asarray(vol.dims) - 1.0) * np.asarray(vol.spacing)
            tr, tri = np.eye(4, dtype=f32), np.eye(4, dtype=f32)
            tr[:3, 3], tri[:3, 3] = c, -c
            nominal = (tr @ rot @ tri @ nominal).astype(f32)
        poses = synth.pose_population(vol, nominal, 3, sigma=(4, 4, 4, 3, 3, 3))
        ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
        assert mask.sum() > 0.3 * mask.size
        rc = _make_rc(ctx, vol, [cam], 3, "pax")
        rc.set_layout_order(variant << 1)
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.compute()
        got = rc.raw_host_pixel_buf()
        gmask, gsteps, gS = rc.ray_info()
        np.testing.assert_array_equal(gmask, mask)
        np.testing.assert_array_equal(gsteps, steps)
        _check_drr(got, ref, mask)
        rc.close()


@pytest.mark.parametrize("frame_type", [0, 1, 2])
@pytest.mark.parametrize("det", [(40, 48), (33, 17), (16, 16), (1, 70)])
def test_drr_frame_types_and_odd_detectors(ctx, xo, small_scene, frame_type, det):
    vol, _, nominal = small_scene
    cam = CameraModel(coord_frame_type=frame_type).setup(400.0, det[0], det[1], 3.2, 3.0)
    if frame_type == 0:
        nominal = (nominal @ np.diag([1, -1, -1, 1]).astype(f32)).astype(f32)
    if frame_type == 2:
        shift = np.eye(4, dtype=f32)
        shift[2, 3] = -400.0
        nominal = (nominal @ shift).astype(f32)
        cam.pinhole_pt = np.array([0, 0, 400.0], dtype=f32)
    poses = synth.pose_population(vol, nominal, 2, sigma=(15, 15, 15, 10, 10, 10))
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    rc = _make_rc(ctx, vol, [cam], 2)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    gmask, gsteps, gS = rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    _check_drr(rc.raw_host_pixel_buf(), ref, mask)


def test_oblique_direction_and_camera_extrinsics(ctx, xo):
    base = synth.make_volume(40, 36, 44, spacing=(0.7, 1.2, 0.9))
    D = xreg_b200.exp_se3([0.3, -0.4, 0.2, 0, 0, 0])[:3, :3].astype(np.float64)
    vol = Volume(base.data, spacing=base.spacing, origin=(5.0, -3.0, 11.0), direction=D)
    K = CameraModel().setup(350.0, 50, 60, 1.1, 1.3).intrins
    K[0, 1] = 0.02  # skew: full 3x3 inverse intrinsics
    E = xreg_b200.exp_se3([0.1, 0.2, -0.1, 5.0, -4.0, 3.0])
    cam = CameraModel().setup_intrins_extrins(K, E, 50, 60, 1.1, 1.3)
    # put the volume centre 200 mm in front of the camera-world origin
    centre = (vol.idx_to_phys().reshape(3, 4) @ np.array([19.5, 17.5, 21.5, 1.0])).astype(f32)
    T = np.eye(4, dtype=f32)
    T[:3, 3] = centre + np.array([0, 0, 200.0], dtype=f32)
    poses = np.stack([(T @ xreg_b200.exp_se3(x)).astype(f32) for x in
                      ([0, 0, 0, 0, 0, 0], [0.02, -0.03, 0.5, 3, -2, 10], [0.05, 0.04, -1.0, -6, 5, -20])])
    poses = np.stack([(p @ cam.extrins).astype(f32) for p in poses])  # camera-world -> volume
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    assert S > 1000
    rc = _make_rc(ctx, vol, [cam], 3)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    gmask, gsteps, _ = rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    _check_drr(rc.raw_host_pixel_buf(), ref, mask)


def test_grazing_missing_and_axis_parallel_rays(ctx, xo):
    vol = Volume(np.ones((16, 16, 16), dtype=f32), spacing=(1, 1, 1), origin=(0, 0, 0))
    cam = CameraModel().setup(1000.0, 64, 64, 0.2, 0.2)
    T = np.eye(4, dtype=f32)
    T[:3, 3] = [15.0 + 500.0 / np.sqrt(2.0), 7.5, 15.0 - 500.0 / np.sqrt(2.0)]
    graze = (T @ xreg_b200.exp_se3([0.0, 3 * np.pi / 4, 0.0, 0, 0, 0])).astype(f32)
    behind = np.eye(4, dtype=f32)
    behind[:3, 3] = [7.5, 7.5, -50.0]
    beyond = np.eye(4, dtype=f32)
    beyond[:3, 3] = [7.5, 7.5, 1600.0]
    poses = np.stack([graze, behind, beyond])
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    assert mask[0].any() and not mask[0].all() and not mask[1:].any()
    rc = _make_rc(ctx, vol, [cam], 3)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    gmask, gsteps, gS = rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    assert gS == S
    _check_drr(rc.raw_host_pixel_buf(), ref, mask)
    # axis-parallel centre ray exactly on / just outside the x = 0 face (power-of-two camera)
    cam2 = CameraModel().setup(512.0, 17, 17, 1.0, 1.0)
    on_face = np.eye(4, dtype=f32)
    on_face[:3, 3] = [0.0, 3.0, 100.0]
    outside = np.eye(4, dtype=f32)
    outside[:3, 3] = [-0.5, 3.0, 100.0]
    p2 = np.stack([on_face, outside])
    ref2, mask2, steps2, _ = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam2)], to12(p2), want_info=True)
    rc2 = _make_rc(ctx, vol, [cam2], 2)
    rc2.set_xforms_cam_to_itk_phys(list(p2))
    rc2.compute()
    gmask2, gsteps2, _ = rc2.ray_info()
    assert gmask2[0, 8, 8] == 1 and gmask2[1, 8, 8] == 0
    np.testing.assert_array_equal(gmask2, mask2)
    np.testing.assert_array_equal(gsteps2, steps2)
    _check_drr(rc2.raw_host_pixel_buf(), ref2, mask2)


def test_mask_bit_exact_random_poses(ctx, xo, small_scene):
    """many random poses incl. ones where the beam only partly covers the volume"""
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, 24, sigma=(40, 40, 40, 40, 40, 60), seed=7)
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    assert 0.05 < mask.mean() < 0.999
    rc = _make_rc(ctx, vol, [cam], 24)
    rc.set_poses_array(to12(poses))
    rc.compute()
    gmask, gsteps, gS = rc.ray_info()
    assert int((gmask != mask).sum()) == 0
    assert int((gsteps != steps).sum()) == 0
    _check_drr(rc.raw_host_pixel_buf(), ref, mask)


def test_step_size_store_methods_bg_max_kernel(ctx, xo, small_scene):
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, 3)
    cams = [xo.cam_struct(cam)]
    rc = _make_rc(ctx, vol, [cam], 3)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    # half-voxel-ish step
    rc.set_ray_step_size(0.45)
    rc.compute()
    ref, mask, steps, _ = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), step_size=0.45, want_info=True)
    np.testing.assert_array_equal(rc.ray_info()[1], steps)
    _check_drr(rc.raw_host_pixel_buf(), ref, mask)
    rc.set_ray_step_size(1.0)
    # REPLACE with a default background value
    rc.set_default_bg_pixel_val(1.5)
    rc.compute()
    buf = np.zeros_like(ref)
    xo.pre_compute(buf, np.zeros(3, np.uint32), None, 0, 1.5)
    xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), buf=buf)
    np.testing.assert_allclose(rc.raw_host_pixel_buf(), buf, rtol=DRR_REL_TOL)
    rc.set_default_bg_pixel_val(0.0)
    # ACCUM: second compute adds on top (multi-object usage, xregIntensity2D3DRegi.cpp:598,628)
    rc.compute()
    once = rc.raw_host_pixel_buf()
    rc.use_proj_store_accum_method()
    rc.compute()
    np.testing.assert_allclose(rc.raw_host_pixel_buf(), 2 * once, rtol=1e-6)
    rc.use_proj_store_replace_method()
    # background projections
    bg = np.random.default_rng(5).random((cam.num_det_rows, cam.num_det_cols)).astype(f32)
    rc.set_bg_proj(bg)
    rc.compute()
    np.testing.assert_allclose(rc.raw_host_pixel_buf(), bg[None] + once, rtol=1e-6)
    rc.set_use_bg_projs(False)
    # max kernel
    rc.set_kernel_id(rc.kRAY_CAST_LINE_INT_MAX_KERNEL)
    rc.compute()
    mref = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), kernel_id=1)
    np.testing.assert_allclose(rc.raw_host_pixel_buf(), mref, rtol=DRR_REL_TOL, atol=0)


def test_multi_view_distribution_and_num_projs_shrink(ctx, xo, small_scene):
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    pop = synth.pose_population(vol, nominal, 3)
    rc = _make_rc(ctx, vol, [cam, cam2], 6)
    rc.distribute_xforms_among_cam_models(list(pop))
    assert rc.camera_model_proj_associations() == [0, 0, 0, 1, 1, 1]
    rc.compute()
    got = rc.raw_host_pixel_buf()
    poses, idx = xo.distribute_xforms(to12(pop), 2)
    ref, mask, _, _ = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam), xo.cam_struct(cam2)], poses, cam_idx=idx,
                             want_info=True)
    _check_drr(got, ref, mask)
    # set_num_projs after allocation with a smaller count
    rc.set_num_projs(2)
    rc.distribute_xform_among_cam_models(pop[1])
    rc.compute()
    small = rc.raw_host_pixel_buf()
    assert small.shape[0] == 2
    np.testing.assert_array_equal(small[0], got[1])
    np.testing.assert_array_equal(small[1], got[4])
    with pytest.raises(xreg_b200.XregError):
        rc.set_num_projs(7)


def test_error_behaviour(ctx, small_scene):
    vol, cam, _ = small_scene
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    with pytest.raises(xreg_b200.XregError):
        rc.compute()  # not allocated (xregRayCastLineIntCPU.cpp:296)
    rc.set_volume(vol)
    bad = CameraModel().setup(400.0, cam.num_det_rows + 2, cam.num_det_cols, 1.0, 1.0)
    with pytest.raises(xreg_b200.XregError):
        rc.set_camera_models([cam, bad])  # mixed detector sizes (xregRayCastBaseCPU.cpp:60-70)
    rc.set_camera_model(cam)
    rc.set_num_projs(1)
    rc.allocate_resources()
    rc.use_bspline_interp()
    with pytest.raises(xreg_b200.UnsupportedOperationException):
        rc.compute()
    rc.use_linear_interp()
    rc.compute()
    with pytest.raises(xreg_b200.XregError):
        rc.compute(3)  # no such volume
    assert rc.max_num_projs_possible() > 1000


def test_multiple_volumes_accumulate(ctx, xo, small_scene):
    vol, cam, nominal = small_scene
    vol2 = synth.make_volume(32, 40, 36, spacing=(1.2, 1.0, 1.1), seed=99)
    poses = synth.pose_population(vol, nominal, 2)
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volumes([vol, vol2])
    rc.set_camera_model(cam)
    rc.set_num_projs(2)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute(0)
    rc.use_proj_store_accum_method()
    rc.compute(1)
    cams = [xo.cam_struct(cam)]
    buf = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses))
    xo.drr(vol2.data, vol2.idx_to_phys(), cams, to12(poses), buf=buf)
    np.testing.assert_allclose(rc.raw_host_pixel_buf(), buf, rtol=DRR_REL_TOL, atol=1e-7)


def test_hu_volume_converted_on_device_is_bit_identical(ctx, xo, small_scene):
    """xrc_rc_set_volumes_hu == HUToLinAtt on the host followed by set_volumes, bit for bit (f64 arithmetic in the
    reference's order on both sides)."""
    vol, cam, nominal = small_scene
    rng = np.random.default_rng(8)
    hu = (vol.data * f32(40000.0) - f32(1000.0) + rng.normal(0, 30, vol.data.shape).astype(f32)).astype(f32)  # -1000 .. ~1500 HU
    poses = synth.pose_population(vol, nominal, 3)
    outs = []
    for hu_lower in (-1000.0, -300.0):
        vol_hu = Volume(hu, vol.spacing, vol.origin, vol.direction)
        vol_att = Volume(xo.hu_to_lin_att(hu, hu_lower), vol.spacing, vol.origin, vol.direction)
        rc_a = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc_a.set_volumes_hu([vol_hu], hu_lower)
        rc_b = _make_rc(ctx, vol_att, [cam], 3)
        rc_a.set_camera_model(cam)
        rc_a.set_num_projs(3)
        rc_a.allocate_resources()
        for rc in (rc_a, rc_b):
            rc.set_xforms_cam_to_itk_phys(list(poses))
            rc.compute()
        a, b = rc_a.raw_host_pixel_buf(), rc_b.raw_host_pixel_buf()
        assert a.tobytes() == b.tobytes() and a.max() > 0
        outs.append(a.copy())
        rc_a.close()
        rc_b.close()
    assert outs[1].sum() < outs[0].sum()   # the higher threshold removes attenuation
