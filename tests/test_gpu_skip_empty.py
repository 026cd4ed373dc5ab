"""Empty-space trimming of the DRR sum kernel (drr.cu): skipping samples that a block map proves to be
zero must not change a single bit of any projection, whatever the volume content, view or step size,
and the result must still match the CPU oracle (xregRayCastLineIntCPU.cpp:270-279 sums every sample)."""
import numpy as np
import pytest

import xreg_b200
from xreg_b200 import synth
from xreg_b200.geometry import CameraModel, Volume, to12

pytestmark = pytest.mark.gpu
f32 = np.float32
DRR_REL_TOL = 1.0e-4


def _rc(ctx, vol, cam, n, step=1.0):
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_ray_step_size(step)
    rc.set_num_projs(n)
    rc.allocate_resources()
    return rc


def _on_off(ctx, vol, cam, poses, step=1.0):
    rc = _rc(ctx, vol, cam, len(poses), step)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.set_skip_empty(True)
    rc.compute()
    on = rc.raw_host_pixel_buf().copy()
    f_on = rc.fetched_samples()
    rc.set_skip_empty(False)
    rc.compute()
    off = rc.raw_host_pixel_buf().copy()
    f_off = rc.fetched_samples()
    S = rc.ray_info(counts_only=True)[2]
    rc.close()
    assert f_off == S                       # without trimming every algorithmic sample is fetched
    assert f_on <= S
    assert on.tobytes() == off.tobytes()    # bit-identical projections
    return on, f_on, S


def _modes(ctx, vol, cam, poses, step=1.0):
    """projections and fetched samples for trimming off / ends only / ends + interior gaps"""
    rc = _rc(ctx, vol, cam, len(poses), step)
    rc.set_xforms_cam_to_itk_phys(list(poses))
    out = {}
    for mode in (0, 1, 3):
        rc.set_skip_empty(mode)
        rc.compute()
        out[mode] = (rc.raw_host_pixel_buf().copy(), rc.fetched_samples())
    rc.close()
    return out


def _far_apart():
    """two slabs 64 voxels apart along x (the map's dilated hulls leave 6 clear blocks between them): a view along x
    sends most rays through both"""
    rng = np.random.default_rng(11)
    d = np.zeros((48, 48, 112), f32)
    d[4:44, 4:44, 8:24] = rng.uniform(0.01, 0.05, (40, 40, 16)).astype(f32)      # two slabs across the view along x
    d[6:42, 2:46, 88:104] = rng.uniform(0.01, 0.05, (36, 44, 16)).astype(f32)
    return Volume(d, spacing=(1.0, 1.0, 1.0), origin=(-55.5, -23.5, -23.5), direction=np.eye(3))


@pytest.mark.parametrize("name", ["far_apart", "phantom", "two_blobs", "single_voxels", "all_zero", "dense"])
@pytest.mark.parametrize("det", [(72, 88, 4), (200, 232, 9)])   # few CTAs (deep pipelines) / the throughput variant
def test_interior_gap_skipping_is_bit_exact(ctx, name, det):
    """xrc_rc_set_skip_empty(rc, 3): the warp also skips runs of empty samples inside a ray's trimmed range (the air
    between two structures).  Same bits in every mode; fewer samples fetched where there is a gap to skip."""
    vol = _far_apart() if name == "far_apart" else _volumes()[name]
    rows, cols, n = det
    cam = CameraModel().setup(420.0, rows, cols, 1.5 * 72 / rows, 1.4 * 88 / cols)
    for view in (0.0, 90.0, 40.0):
        poses = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=view), n,
                                      sigma=(12, 12, 12, 8, 8, 8))
        m = _modes(ctx, vol, cam, poses)
        for mode in (1, 3):
            assert m[mode][0].tobytes() == m[0][0].tobytes(), (name, view, mode)
        assert m[3][1] <= m[1][1] <= m[0][1]
        if name == "far_apart" and view == 90.0:
            assert m[3][1] < 0.85 * m[1][1]         # looking along x, through both structures: the gap is skipped
        if name == "dense":
            assert m[3][1] == m[0][1]


def _volumes():
    rng = np.random.default_rng(7)
    out = {}
    out["phantom"] = synth.make_volume(72, 64, 56, spacing=(0.9, 1.1, 1.3))
    # bone-mask-like: two separated blobs, zero elsewhere (interior gap between them)
    d = np.zeros((56, 64, 72), f32)
    d[10:30, 8:28, 6:30] = rng.uniform(0.01, 0.05, (20, 20, 24)).astype(f32)
    d[30:50, 40:60, 44:70] = rng.uniform(0.01, 0.05, (20, 20, 26)).astype(f32)
    out["two_blobs"] = Volume(d, spacing=(1.0, 1.0, 1.0), origin=(-35.5, -31.5, -27.5), direction=np.eye(3))
    # a single non-zero voxel at a block corner, and one on the volume face
    d = np.zeros((40, 48, 56), f32)
    d[16, 24, 32] = 1.0
    d[0, 0, 55] = 2.0
    d[39, 47, 0] = -3.0
    out["single_voxels"] = Volume(d, spacing=(1.2, 0.8, 1.0), origin=(-33.0, -18.8, -19.5), direction=np.eye(3))
    # all zero, incl. negative zeros
    d = np.zeros((24, 24, 24), f32)
    d[::2] = -0.0
    out["all_zero"] = Volume(d, spacing=(1.0, 1.0, 1.0), origin=(-11.5, -11.5, -11.5), direction=np.eye(3))
    # dense (no air at all): nothing to trim
    d = rng.uniform(0.01, 0.02, (32, 40, 48)).astype(f32)
    out["dense"] = Volume(d, spacing=(1.0, 1.0, 1.0), origin=(-23.5, -19.5, -15.5), direction=np.eye(3))
    return out


@pytest.mark.parametrize("name", ["phantom", "two_blobs", "single_voxels", "all_zero", "dense"])
@pytest.mark.parametrize("view_deg", [0.0, 90.0, 40.0])
def test_trimming_is_bit_exact_and_matches_oracle(ctx, xo, name, view_deg):
    vol = _volumes()[name]
    cam = CameraModel().setup(420.0, 72, 88, 1.5, 1.4)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=view_deg)
    poses = synth.pose_population(vol, nominal, 4, sigma=(12, 12, 12, 8, 8, 8))
    got, f_on, S = _on_off(ctx, vol, cam, poses)
    ref, mask, steps, So = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), want_info=True)
    assert So == S
    assert np.all(got[mask == 0] == 0)
    sel = (mask == 1) & (np.abs(ref) > 0)
    if sel.any():
        assert (np.abs(got[sel] - ref[sel]) / np.abs(ref[sel])).max() <= DRR_REL_TOL
    assert np.all(got[(mask == 1) & (ref == 0)] == 0)
    if name == "all_zero":
        assert f_on < 0.1 * S           # only rays that take the clamped loop (no drift proof) are left untrimmed
    if name == "dense":
        assert f_on == S
    if name in ("two_blobs", "single_voxels"):
        assert f_on < 0.85 * S


@pytest.mark.parametrize("step", [0.25, 0.5, 2.5, 7.0])
def test_trimming_with_other_step_sizes(ctx, xo, step):
    """the number of samples one map bit vouches for depends on the step (m = floor(6 / max axis step));
    a per-axis step longer than the reach turns trimming off for that ray"""
    vol = _volumes()["two_blobs"]
    cam = CameraModel().setup(420.0, 40, 48, 2.8, 2.6)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0, view_rot_deg=20.0)
    poses = synth.pose_population(vol, nominal, 3, sigma=(10, 10, 10, 6, 6, 6))
    got, f_on, S = _on_off(ctx, vol, cam, poses, step=step)
    ref, mask, _, _ = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses), step_size=step, want_info=True)
    sel = (mask == 1) & (np.abs(ref) > 0)
    assert (np.abs(got[sel] - ref[sel]) / np.abs(ref[sel])).max() <= DRR_REL_TOL
    if step >= 7.0:
        assert f_on > 0.9 * S   # the per-axis step exceeds the 6-voxel reach for (almost) every ray


def test_trimming_with_scaled_pose_and_anisotropic_direction(ctx):
    """non-rigid cam->phys matrices change the index-space step per ray: the reach is evaluated per ray"""
    base = _volumes()["two_blobs"]
    D = xreg_b200.exp_se3([0.3, -0.4, 0.2, 0, 0, 0])[:3, :3].astype(np.float64)
    vol = Volume(base.data, spacing=(0.5, 1.7, 1.0), origin=(3.0, -2.0, 4.0), direction=D)
    cam = CameraModel().setup(420.0, 40, 48, 2.8, 2.6)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0)
    poses = synth.pose_population(vol, nominal, 3, sigma=(10, 10, 10, 6, 6, 6))
    poses[1, :3, :3] *= 1.7   # not rigid
    poses[2, :3, :3] *= 0.4
    _on_off(ctx, vol, cam, poses)


def test_trimming_accum_and_background(ctx):
    vol = _volumes()["two_blobs"]
    cam = CameraModel().setup(420.0, 40, 48, 2.8, 2.6)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0)
    poses = synth.pose_population(vol, nominal, 2)
    outs = []
    for skip in (True, False):
        rc = _rc(ctx, vol, cam, 2)
        rc.set_skip_empty(skip)
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.set_default_bg_pixel_val(0.25)
        rc.compute()
        rc.use_proj_store_accum_method()
        rc.compute()
        outs.append(rc.raw_host_pixel_buf().copy())
        rc.close()
    assert outs[0].tobytes() == outs[1].tobytes()


def test_max_kernel_never_trims(ctx):
    vol = _volumes()["single_voxels"]   # holds a negative voxel: max over zeros differs from max over nothing
    cam = CameraModel().setup(420.0, 40, 48, 2.8, 2.6)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0)
    poses = synth.pose_population(vol, nominal, 2)
    outs = []
    for skip in (True, False):
        rc = _rc(ctx, vol, cam, 2)
        rc.set_kernel_id(1)  # XRC_KERNEL_MAX
        rc.set_skip_empty(skip)
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.compute()
        outs.append(rc.raw_host_pixel_buf().copy())
        S = rc.ray_info(counts_only=True)[2]
        assert rc.fetched_samples() == S
        rc.close()
    assert outs[0].tobytes() == outs[1].tobytes()


def test_trimming_full_size_c2_population(ctx):
    """BASELINE config C2 geometry (512x512x400 CT, 480^2 detector): trimmed and untrimmed launches agree
    bit for bit over a population, and the trimmed one fetches measurably fewer samples."""
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    cam = synth.make_camera(480)
    nominal = synth.nominal_pose(vol)
    poses = synth.pose_population(vol, nominal, 6)
    _, f_on, S = _on_off(ctx, vol, cam, poses)
    assert f_on < 0.95 * S
