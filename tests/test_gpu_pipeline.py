"""End-to-end GPU tests through the C ABI: the fused batch evaluation (obj_fn), and
BASELINE.json's full-size configuration (512x512x400 CT, 480x480 detector, patch
gradient-NCC, population 100) checked against the oracle on a bounded sample plus
size-independent properties."""
import numpy as np
import pytest

import xreg_b200
from xreg_b200 import regi, synth
from xreg_b200.geometry import CameraModel, to12

pytestmark = pytest.mark.gpu
f32 = np.float32


def test_eval_batch_equals_separate_calls_multi_view(ctx, xo, small_scene):
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    cams = [cam, cam2]
    pop = synth.pose_population(vol, nominal, 6)
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    for metric, ofn in (("patch-grad-ncc", lambda f, d: xo.patch_grad_ncc(f, d, xo.patch_opts(radius=6))),
                        ("grad-ncc", lambda f, d: xo.grad_ncc(f, d)),
                        ("ncc", lambda f, d: xo.ncc(f, d))):
        fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=metric, max_pop=6, patch_radius=6)
        got = fn(pop)
        poses, idx = xo.distribute_xforms(to12(pop), 2)
        drr = xo.drr(vol.data, vol.idx_to_phys(), xcams, poses, cam_idx=idx)
        ref = xo.combine_mean(np.stack([ofn(fixed[0], drr[:6]), ofn(fixed[1], drr[6:])]))
        assert np.max(np.abs(got - ref)) <= 1e-5, metric
        assert int(np.argmin(got)) == 0
        # smaller populations re-use the allocation (BOBYQA: population 1)
        one = fn(pop[2:3])
        assert abs(one[0] - got[2]) <= 1e-7
        three = fn(pop[[4, 1, 3]])
        np.testing.assert_array_equal(three, got[[4, 1, 3]])


def test_objective_from_se3_parameters(ctx, small_scene):
    """xrc_obj_fn_se3: pose = pre * ExpSE3(x) * post composed inside the library equals the same poses
    composed on the host and passed to xrc_obj_fn."""
    from xreg_b200.geometry import exp_se3

    vol, cam, nominal = small_scene
    c = np.asarray(vol.origin) + 0.5 * (np.asarray(vol.dims) - 1.0) * np.asarray(vol.spacing)
    pre, post = np.eye(4, dtype=f32), np.eye(4, dtype=f32)
    pre[:3, 3] = c
    post[:3, 3] = -c
    post = (post @ nominal).astype(f32)   # pre * delta * post = C delta C^-1 nominal
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(0, 0.05, (7, 3)), rng.normal(0, 4.0, (7, 3))], axis=1).astype(f32)
    x[0] = 0
    fixed = np.ones((cam.num_det_rows, cam.num_det_cols), f32)
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="ncc", max_pop=7)
    rc0 = fn.rc
    poses = np.stack([(pre @ exp_se3(xi) @ post).astype(f32) for xi in x])
    a = fn(poses)
    fixed = synth.add_noise(rc0.proj(0))
    fn2 = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="grad-ncc", max_pop=7)
    a = fn2(poses)
    b = fn2.eval_se3(x, pre, post)
    assert np.max(np.abs(a - b)) <= 1e-5
    assert int(np.argmin(b)) == 0
    one = fn2.eval_se3(x[3:4], pre, post)
    assert abs(one[0] - b[3]) <= 1e-7


def _device_lists():
    import torch

    lists = [[0, 0], [0, 0, 0]]          # several contexts / streams on one GPU: always available
    if torch.cuda.device_count() >= 2:
        lists.append(list(range(min(torch.cuda.device_count(), 8))))
    return lists


@pytest.mark.parametrize("metric", ["patch-grad-ncc", "grad-ncc"])
def test_multi_device_objective_equals_single_device(ctx, xo, small_scene, metric):
    """xrc_obj_fn_multi: the population split over several device replicas from one host thread gives the
    single-device values bit for bit, for even / uneven chunks, fewer poses than devices, and two views."""
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    cams = [cam, cam2]
    pop = synth.pose_population(vol, nominal, 11)
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    single = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=metric, max_pop=11, patch_radius=6)
    ref = single(pop)
    ref_pv = np.stack([sm.sim_vals()[:11] for sm in single.sims])
    for devices in _device_lists():
        multi = regi.MultiDeviceObjFn(devices, vol, cams, fixed, max_pop=11, metric=metric, patch_radius=6)
        got = multi(pop)
        np.testing.assert_array_equal(got, ref)
        np.testing.assert_array_equal(multi.per_view, ref_pv)
        np.testing.assert_array_equal(multi(pop[3:10]), ref[3:10])       # 7 poses: uneven chunks
        np.testing.assert_array_equal(multi(pop[5:6]), ref[5:6])         # 1 pose: the other devices idle
        np.testing.assert_array_equal(multi(pop[::-1]), ref[::-1])
        with pytest.raises(xreg_b200.XregError):
            multi(np.concatenate([pop, pop]))                            # over capacity
        multi.close()


def test_multi_device_objective_shards_views(ctx, xo, small_scene):
    """the camera-major (view, pose) list is cut into contiguous chunks that may straddle views,
    so a small multi-view population (BOBYQA: one pose, three views) still spreads over the devices.  Values are the
    single-device ones bit for bit, whatever the split."""
    vol, cam, nominal = small_scene
    cams = [cam, CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4),
            CameraModel().setup(420.0, cam.num_det_rows, cam.num_det_cols, 1.5, 1.6)]
    pop = synth.pose_population(vol, nominal, 7)
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    single = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric="grad-ncc", max_pop=7)
    ref = single(pop)
    ref_pv = np.stack([sm.sim_vals()[:7] for sm in single.sims])
    lists = _device_lists() + [[0] * 4, [0] * 5]
    for devices in lists:
        multi = regi.MultiDeviceObjFn(devices, vol, cams, fixed, max_pop=7, metric="grad-ncc")
        for sel in (slice(0, 7), slice(2, 3), slice(1, 3), slice(0, 5), slice(6, 7)):
            got = multi(pop[sel])
            np.testing.assert_array_equal(got, ref[sel])
            np.testing.assert_array_equal(multi.per_view, ref_pv[:, sel])
        # the replicas stay usable on their own after the library re-sized them
        np.testing.assert_array_equal(multi.replicas[0](pop[:1]), ref[:1])
        np.testing.assert_array_equal(multi(pop[3:6]), ref[3:6])
        multi.close()
    # oracle: the three-view mean of one pose
    d = [xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[2:3]))[0] for xc in xcams]
    want = xo.combine_mean(np.stack([xo.grad_ncc(fixed[v], d[v][None], gauss_width=5) for v in range(3)]))
    assert abs(float(want[0]) - float(ref[2])) <= 1e-5


def test_objective_units_equal_the_full_objective(ctx, xo, small_scene):
    """xrc_obj_fn_units (one rank's part of a (view, pose)-sharded objective): any contiguous range of the camera-major
    unit list reproduces the full call's per-view values bit for bit; ShardedViewObjFn on one rank = the full objective."""
    vol, cam, nominal = small_scene
    cams = [cam, CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4),
            CameraModel().setup(420.0, cam.num_det_rows, cam.num_det_cols, 1.5, 1.6)]
    pop = synth.pose_population(vol, nominal, 5)
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pop[:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric="patch-grad-ncc", max_pop=5, patch_radius=6)
    ref = fn(pop)
    flat = np.stack([sm.sim_vals()[:5] for sm in fn.sims]).reshape(-1)
    for first, count in [(0, 15), (0, 1), (4, 3), (5, 5), (7, 8), (14, 1), (3, 0)]:
        np.testing.assert_array_equal(fn.eval_units(pop, first, count), flat[first:first + count])
    np.testing.assert_array_equal(fn(pop), ref)                      # the object still works as a whole afterwards
    one = fn.eval_units(pop[2:3], 1, 2)                              # one pose: views 1 and 2
    np.testing.assert_array_equal(one, flat.reshape(3, 5)[1:, 2])
    sharded = regi.ShardedViewObjFn(fn.eval_units, 3)
    np.testing.assert_array_equal(sharded(pop), ref)
    with pytest.raises(xreg_b200.XregError):
        fn.eval_units(pop, 10, 6)                                    # beyond the 15 units
    fn.close()


@pytest.mark.parametrize("n_poses", [1, 5])
def test_multi_object_objective_matches_oracle(ctx, xo, small_scene, n_poses):
    """xrc_obj_fn_objects: two moving volumes with their own pose populations accumulated into the same projections
    (first REPLACE, then ACCUM; xregIntensity2D3DRegi.cpp:594-629), with and without a static background projection."""
    vol, cam, nominal = small_scene
    rng = np.random.default_rng(11)
    d2 = np.zeros((30, 36, 40), f32)
    d2[6:22, 8:30, 10:34] = rng.uniform(0.02, 0.06, (16, 22, 24)).astype(f32)
    vol2 = xreg_b200.Volume(d2, spacing=(1.2, 0.9, 1.1), origin=(-20.0, -12.0, -10.0), direction=np.eye(3))
    xcam = [xo.cam_struct(cam)]
    pop_a = synth.pose_population(vol, nominal, n_poses, seed=1)
    pop_b = synth.pose_population(vol, nominal, n_poses, seed=2, sigma=(8, 8, 8, 6, 6, 6))
    fixed = synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop_a[:1]))[0])
    bg = np.abs(rng.normal(0.5, 0.1, fixed.shape)).astype(f32)
    fn = regi.Intensity2D3DObjFn(ctx, [vol, vol2], [cam], [fixed], metric="grad-ncc", max_pop=5)
    fn.rc.set_bg_projs([bg], use_bg_projs=True)   # uploads the image
    fn.rc.set_use_bg_projs(False)                 # ... but plain calls do not use it
    for use_bg in (False, True):
        got = fn.eval_objects([pop_a, pop_b], use_bg_projs=use_bg)
        buf = np.repeat(bg[None], n_poses, axis=0).copy() if use_bg else np.zeros((n_poses,) + fixed.shape, f32)
        xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop_a), buf=buf)
        xo.drr(vol2.data, vol2.idx_to_phys(), xcam, to12(pop_b), buf=buf)
        drr = fn.rc.raw_host_pixel_buf()[:n_poses]
        sel = buf > 1e-3 * buf.max()
        assert np.max(np.abs(drr[sel] - buf[sel]) / buf[sel]) <= 1e-4
        assert np.max(np.abs(got - xo.grad_ncc(fixed, buf))) <= 1e-5
    # the ray caster's own settings are untouched: a plain call still replaces and ignores the background
    plain = fn(pop_a)
    ref = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop_a))
    assert np.max(np.abs(plain - xo.grad_ncc(fixed, ref))) <= 1e-5
    # swapped order / explicit volume indices: object list [vol2, vol] gives the same projections up to rounding
    swapped = fn.eval_objects([pop_b, pop_a], vol_inds=[1, 0])
    assert np.max(np.abs(swapped - fn.eval_objects([pop_a, pop_b]))) <= 1e-5


def test_launch_counter_counts_our_kernels(ctx, small_scene):
    vol, cam, nominal = small_scene
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [np.ones((cam.num_det_rows, cam.num_det_cols), f32)], metric="patch-grad-ncc",
                                 max_pop=2, patch_radius=5)
    pop = synth.pose_population(vol, nominal, 2)
    fn(pop)                                        # the first call also builds the volume stack these poses need
    before = xreg_b200.launch_count()
    fn(pop)
    assert xreg_b200.launch_count() - before == 5  # drr + grad + patch + sequential patch sum + finalize


@pytest.fixture(scope="module")
def full_scene():
    vol = synth.make_volume(512, 512, 400, spacing=(0.8, 0.8, 1.0))
    cam = synth.make_camera(480)
    nominal = synth.nominal_pose(vol)
    pop = synth.pose_population(vol, nominal, 100)
    return vol, cam, nominal, pop


def test_full_size_config_parity_sample_and_properties(ctx, xo, full_scene):
    vol, cam, nominal, pop = full_scene
    radius = synth.patch_radius_for(480)
    assert radius == 13
    xcam = [xo.cam_struct(cam)]
    # bounded oracle sample: 8 of the 100 poses, full detector (about 2 s of the oracle on the GPU box's host cores)
    sample = [0, 57, 13, 99, 31, 76, 42, 5]
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop[sample]), want_info=True)
    fixed = synth.add_noise(ref[0])
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="patch-grad-ncc", max_pop=100, patch_radius=radius)
    sims = fn(pop)
    assert sims.shape == (100,) and np.all(np.isfinite(sims))
    assert int(np.argmin(sims)) == 0
    # DRR parity on the sampled poses (projection buffer still holds the batch)
    for k, p in enumerate(sample):
        got = fn.rc.proj(p)
        np.testing.assert_array_equal(got[mask[k] == 0], ref[k][mask[k] == 0])
        sel = (mask[k] == 1) & (ref[k] > 0)
        assert np.max(np.abs(got[sel] - ref[k][sel]) / ref[k][sel]) <= 1e-4
    # similarity parity on the sampled poses
    osim = xo.patch_grad_ncc(fixed, ref, xo.patch_opts(radius=radius))
    assert np.max(np.abs(sims[sample] - osim)) <= 1e-5
    # masks / step counts bit exact at full size
    fn.rc.set_num_projs(len(sample))
    fn.rc.set_poses_array(to12(pop[sample]))
    gmask, gsteps, gS = fn.rc.ray_info()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    assert gS == S == int(steps.sum())
    fn._cur_pop = -1  # force re-binding after the manual set_num_projs above
    # batch-order invariance and sub-batch idempotence (bitwise)
    perm = np.random.default_rng(0).permutation(100)
    np.testing.assert_array_equal(fn(pop[perm]), sims[perm])
    np.testing.assert_array_equal(fn(pop[10:23]), sims[10:23])
    np.testing.assert_array_equal(fn(pop), sims)
    # linearity of the line integral in the volume: DRR(2 v) == 2 DRR(v) exactly (power of two)
    d1 = fn.rc.proj(5)
    vol2 = xreg_b200.Volume(vol.data * f32(2), vol.spacing, vol.origin, vol.direction)
    rc2 = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc2.set_volume(vol2)
    rc2.set_camera_model(cam)
    rc2.set_num_projs(1)
    rc2.allocate_resources()
    rc2.set_xforms_cam_to_itk_phys([pop[5]])
    rc2.compute()
    np.testing.assert_array_equal(rc2.proj(0), d1 * f32(2))
