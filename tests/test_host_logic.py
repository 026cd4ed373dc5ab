"""Host-side logic (no GPU): camera models, pose helpers, patch weights, synthetic data."""
import numpy as np
import pytest

from xreg_b200 import synth
from xreg_b200.geometry import (CameraModel, Volume, downsample_camera_model, exp_se3, se3_inv, to12)

f32 = np.float32


def test_camera_naive_setup_matches_oracle(xo):
    for ft in (0, 1, 2):
        cam = CameraModel(coord_frame_type=ft).setup(1020.0, 1536, 1536, 0.194, 0.194)
        ref = xo.cam_setup_naive(1020.0, 1536, 1536, 0.194, 0.194, ft)
        got = xo.cam_struct(cam)
        np.testing.assert_allclose(np.array(got.intrins_inv), np.array(ref.intrins_inv), rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(np.array(got.extrins_inv), np.array(ref.extrins_inv))
        np.testing.assert_array_equal(np.array(got.pinhole), np.array(ref.pinhole))
        assert got.focal_len == ref.focal_len and got.frame_type == ref.frame_type
    assert cam.intrins[0, 2] == 767.5


def test_camera_with_extrinsics_matches_oracle(xo):
    K = CameraModel().setup(1000.0, 64, 48, 0.5, 0.4).intrins
    E = exp_se3([0.2, -0.1, 0.3, 10.0, -20.0, 30.0])
    for ft in (0, 1, 2):
        cam = CameraModel(coord_frame_type=ft).setup_intrins_extrins(K, E, 64, 48, 0.5, 0.4)
        ref = xo.cam_setup(K, E, 64, 48, 0.5, 0.4, ft)
        got = xo.cam_struct(cam)
        np.testing.assert_allclose(np.array(got.extrins_inv), np.array(ref.extrins_inv), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(np.array(got.pinhole), np.array(ref.pinhole), rtol=1e-5, atol=1e-4)
        assert abs(got.focal_len - ref.focal_len) < 1e-3


def test_downsample_camera_model_rules():
    cam = CameraModel().setup(1020.0, 1536, 1536, 0.194, 0.194)
    ds = downsample_camera_model(cam, 0.125)
    assert (ds.num_det_rows, ds.num_det_cols) == (192, 192)
    assert abs(ds.det_col_spacing - 0.194 * 8) < 1e-5
    assert abs(ds.intrins[0, 2] - 767.5 / 8) < 1e-4
    assert abs(ds.focal_len - 1020.0) < 1e-2
    odd = downsample_camera_model(cam, 0.2503, force_even_dims=True)
    assert odd.num_det_rows % 2 == 0


def test_exp_se3_and_inverse():
    T = exp_se3([0.3, -0.2, 0.5, 1.0, 2.0, 3.0])
    np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)
    np.testing.assert_allclose(se3_inv(T) @ T, np.eye(4), atol=1e-5)
    np.testing.assert_array_equal(exp_se3([0, 0, 0, 1, 2, 3])[:3, 3], [1, 2, 3])
    # small-angle limit is a pure rotation about z
    R = exp_se3([0, 0, np.pi / 2, 0, 0, 0])
    np.testing.assert_allclose(R[:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-6)
    assert to12(np.stack([T, T])).shape == (2, 12)


def test_volume_idx_to_phys():
    D = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float64)
    v = Volume(np.zeros((3, 4, 5), f32), spacing=(0.5, 0.75, 2.0), origin=(1, 2, 3), direction=D)
    A = v.idx_to_phys().reshape(3, 4)
    np.testing.assert_allclose(A[:, :3], D * np.array([0.5, 0.75, 2.0])[None, :])
    np.testing.assert_array_equal(A[:, 3], [1, 2, 3])
    assert v.dims == (5, 4, 3)


def test_synthetic_inputs_are_deterministic():
    a = synth.make_volume(32, 28, 24, spacing=(0.8, 0.8, 1.0))
    b = synth.make_volume(32, 28, 24, spacing=(0.8, 0.8, 1.0))
    np.testing.assert_array_equal(a.data, b.data)
    assert a.data[:, :, 0].max() == 0 and 0.05 in a.data and a.data.min() == 0
    T = synth.nominal_pose(a)
    P = synth.pose_population(a, T, 5)
    np.testing.assert_array_equal(P, synth.pose_population(a, T, 5))
    assert synth.patch_radius_for(480) == 13 and synth.patch_radius_for(192) == 5
    assert synth.patch_radius_for(384) == 10 and synth.patch_radius_for(768) == 21


class _FakeMetric:
    """The weight logic of ImgSimMetric2DPatchCommon without a device handle."""

    def __init__(self, fixed, mask):
        from xreg_b200.sim_metrics import ImgSimMetric2DPatchCommon

        self.__class__ = type("W", (ImgSimMetric2DPatchCommon,), {})
        self._init_patch_common()
        self._fixed, self._mask, self._allocated = fixed, mask, False


@pytest.mark.parametrize("radius,stride", [(2, 1), (3, 2)])
def test_patch_weights_match_oracle(xo, radius, stride):
    rng = np.random.default_rng(0)
    fixed = rng.random((23, 29)).astype(f32)
    mask = (rng.random(fixed.shape) > 0.4).astype(np.uint8)
    m = _FakeMetric(fixed, mask)
    m.set_patch_radius(radius)
    m.set_patch_stride(stride)
    o = xo.patch_opts(radius=radius, stride=stride)
    np.testing.assert_allclose(m.compute_weights(), xo.patch_weights(23, 29, o, mask=mask), rtol=2e-6)
    assert m.num_patches() == xo.num_patches(23, 29, radius, stride)
    wi = rng.random(fixed.shape).astype(f32)
    m._wgt_img = wi
    np.testing.assert_allclose(m.compute_weights(), xo.patch_weights(23, 29, o, mask=mask, wgt_img=wi), rtol=2e-6)
    m._wgt_img, m._mask = None, None
    assert m.compute_weights() is None


def test_library_exp_se3_matches_host_model():
    """xrc_exp_se3 (ExpSE3, lib/transforms/xregRigidUtils.cpp:40-85) is host-only code of the product library:
    compare with the numpy restatement, including the small-angle branch."""
    import ctypes as C

    from xreg_b200 import _lib
    from xreg_b200.geometry import exp_se3

    lib = _lib.load()
    rng = np.random.default_rng(11)
    FP = C.POINTER(C.c_float)
    cases = [np.zeros(6), np.array([0, 0, 0, 1.5, -2.0, 3.0]), np.array([1e-20, 0, 0, 1, 2, 3])]
    cases += [np.concatenate([rng.normal(0, 0.4, 3), rng.normal(0, 30, 3)]) for _ in range(20)]
    for x in cases:
        x32 = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros(12, dtype=np.float32)
        lib.xrc_exp_se3(x32.ctypes.data_as(FP), out.ctypes.data_as(FP))
        ref = exp_se3(x32)[:3, :].reshape(12)
        assert np.max(np.abs(out - ref)) <= 2e-5 * max(1.0, float(np.abs(ref).max())), (x, out, ref)
        R = out.reshape(3, 4)[:, :3].astype(np.float64)
        assert np.max(np.abs(R @ R.T - np.eye(3))) < 1e-5


def test_multi_device_partition_covers_every_projection_once():
    """xrc_obj_fn_multi_share (host-only code of the product library): the camera-major (view, pose) list is cut into
    contiguous chunks whose sizes differ by at most one; every projection is owned by exactly one device; with one
    view the chunks are shard_bounds()' (100 poses on 8 devices -> 13 13 13 13 12 12 12 12)."""
    from xreg_b200 import regi

    for n_dev, n_views, n_poses in [(8, 1, 100), (3, 3, 1), (4, 3, 1), (2, 3, 1), (8, 3, 100), (5, 2, 7), (3, 4, 0), (7, 1, 3)]:
        owner = -np.ones((n_views, n_poses), dtype=np.int64)
        sizes = []
        for d in range(n_dev):
            tot = 0
            for v in range(n_views):
                b, e = regi.multi_device_share(n_dev, n_views, n_poses, d, v)
                assert 0 <= b <= e <= n_poses
                assert np.all(owner[v, b:e] == -1)
                owner[v, b:e] = d
                tot += e - b
            sizes.append(tot)
        assert np.all(owner >= 0)
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n_views * n_poses
        flat = owner.reshape(-1)                     # camera-major order: owners are non-decreasing -> contiguous chunks
        assert np.all(np.diff(flat) >= 0)
        if n_views == 1:
            mine = [regi.multi_device_share(n_dev, 1, n_poses, d, 0) for d in range(n_dev)]
            assert [r for r in mine if r[1] > r[0]] == [r for r in regi.shard_bounds(n_poses, n_dev) if r[1] > r[0]]
    assert [regi.multi_device_share(3, 3, 1, d, d) for d in range(3)] == [(0, 1)] * 3      # C4, population 1: a view each
    with pytest.raises(Exception):
        regi.multi_device_share(2, 1, 5, 2, 0)
