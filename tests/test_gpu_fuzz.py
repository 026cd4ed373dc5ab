"""Randomised parity sweep of the DRR path against the CPU oracle: volume sizes down to a single voxel per
axis, anisotropic spacing, oblique direction cosines, random cameras (all three coordinate-frame types),
poses from "looking at the volume" to "camera inside it" and "missing it", step sizes, both line-integral
kernels, REPLACE / ACCUM, sparse and dense contents.  Every case: clip masks and per-ray sample counts
bit-exact, DRR relative error <= 1e-4 (on pixels that are not numerically tiny), trimming on/off bitwise equal."""
import os

import numpy as np
import pytest

import xreg_b200
from xreg_b200 import synth
from xreg_b200.geometry import CameraModel, Volume, to12

pytestmark = pytest.mark.gpu
f32 = np.float32
# XREG_FUZZ_SCALE=k multiplies the number of random cases (long runs recorded under profiles/)
_SCALE = max(1, int(os.environ.get("XREG_FUZZ_SCALE", "1")))


def _scene(seed):
    rng = np.random.default_rng(1000 + seed)
    small = seed % 5 == 0
    dims = [int(rng.integers(1, 4)) if (small and rng.random() < 0.5) else int(rng.integers(2, 41)) for _ in range(3)]
    nx, ny, nz = dims
    kind = seed % 4
    data = rng.uniform(0.0, 0.06, (nz, ny, nx)).astype(f32)
    if kind == 1:      # sparse: a few non-zero blobs
        keep = np.zeros_like(data, dtype=bool)
        for _ in range(3):
            c = [int(rng.integers(0, d)) for d in (nz, ny, nx)]
            r = int(rng.integers(1, 6))
            keep[max(0, c[0] - r):c[0] + r, max(0, c[1] - r):c[1] + r, max(0, c[2] - r):c[2] + r] = True
        data = np.where(keep, data, 0).astype(f32)
    elif kind == 2:    # zero shell around a dense core
        m = int(min(dims) // 4)
        if m > 0:
            core = np.zeros_like(data)
            core[m:nz - m or None, m:ny - m or None, m:nx - m or None] = data[m:nz - m or None, m:ny - m or None, m:nx - m or None]
            data = core
    elif kind == 3:    # negative and large values too
        data = (data - f32(0.03)) * f32(50.0)
    spacing = tuple(float(s) for s in rng.uniform(0.4, 2.5, 3))
    w = rng.normal(0, 0.4, 3)
    D = xreg_b200.exp_se3([w[0], w[1], w[2], 0, 0, 0])[:3, :3].astype(np.float64) if seed % 3 == 0 else np.eye(3)
    origin = tuple(float(o) for o in rng.uniform(-30, 30, 3))
    vol = Volume(data, spacing=spacing, origin=origin, direction=D)
    rows, cols = int(rng.integers(1, 50)), int(rng.integers(1, 50))
    frame = int(rng.integers(0, 3))
    focal = float(rng.uniform(200, 600))
    cam = CameraModel(coord_frame_type=frame).setup(focal, rows, cols, float(rng.uniform(0.5, 4.0)), float(rng.uniform(0.5, 4.0)))
    # volume centre in physical space
    i2p = np.asarray(vol.idx_to_phys(), np.float64).reshape(3, 4)
    centre = i2p @ np.array([(nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0, 1.0])
    extent = float(np.linalg.norm(np.array(dims) * np.array(spacing)))
    poses = []
    for p in range(4):
        x = np.concatenate([rng.normal(0, 0.6, 3), rng.normal(0, 0.15 * extent + 1.0, 3)])
        T = np.eye(4)
        # camera frame: put the volume centre at depth z along the optical axis (sign by frame type); p == 3: inside / behind
        depth = rng.uniform(0.2, 0.7) * focal if p < 3 else rng.uniform(-0.1, 0.1) * extent
        zsign = -1.0 if frame == 1 else 1.0
        T[:3, 3] = centre - np.array([0.0, 0.0, zsign * depth])
        C, Ci = np.eye(4), np.eye(4)
        C[:3, 3], Ci[:3, 3] = centre, -centre
        poses.append((C @ xreg_b200.exp_se3(x) @ Ci @ T).astype(f32))
    step = float(rng.choice([0.25, 0.5, 1.0, 1.0, 2.0, 3.7]))
    kernel_id = int(seed % 7 == 3)
    return vol, cam, np.stack(poses), step, kernel_id, kind


@pytest.mark.parametrize("seed", range(48 * _SCALE))
def test_random_scene_matches_oracle(ctx, xo, seed):
    vol, cam, poses, step, kernel_id, kind = _scene(seed)
    n = poses.shape[0]
    xcam = [xo.cam_struct(cam)]
    ref, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(poses), step_size=step, kernel_id=kernel_id,
                                 want_info=True)
    outs = []
    for skip in (True, False):
        rc = xreg_b200.RayCasterLineIntCUDA(ctx)
        rc.set_volume(vol)
        rc.set_camera_model(cam)
        rc.set_ray_step_size(step)
        rc.set_kernel_id(kernel_id)
        rc.set_skip_empty(skip)
        rc.set_num_projs(n)
        rc.allocate_resources()
        rc.set_xforms_cam_to_itk_phys(list(poses))
        rc.compute()
        got = rc.raw_host_pixel_buf().copy()
        if skip:
            gmask, gsteps, gS = rc.ray_info()
            np.testing.assert_array_equal(gmask, mask)
            np.testing.assert_array_equal(gsteps, steps)
            assert gS == S
            assert rc.fetched_samples() <= S
        # ACCUM on top of the first result: 2x the integral for the sum kernel, unchanged for max
        rc.use_proj_store_accum_method()
        rc.compute()
        twice = rc.raw_host_pixel_buf().copy()
        rc.close()
        if kernel_id == 0:
            np.testing.assert_array_equal(twice, got + got)
        else:
            np.testing.assert_array_equal(twice, np.maximum(got, got))
        outs.append(got)
    assert outs[0].tobytes() == outs[1].tobytes()
    got = outs[0]
    assert np.all(got[mask == 0] == ref[mask == 0])
    scale = float(np.abs(ref).max())
    # signed volumes (kind 3) cancel along the ray: judge the relative error only where little cancelled
    floor = (0.05 if kind == 3 else 1e-3) * scale
    sel = (mask == 1) & (np.abs(ref) > floor) if scale > 0 else np.zeros_like(mask, bool)
    if sel.any():
        assert (np.abs(got[sel] - ref[sel]) / np.abs(ref[sel])).max() <= 1.0e-4
    # numerically tiny pixels: absolute agreement at the rounding level of the largest ones
    assert np.abs(got - ref).max() <= 1.0e-4 * max(scale, 1e-30) + 1e-12


@pytest.mark.parametrize("seed", range(40 * _SCALE))
def test_random_metric_case_matches_oracle(ctx, xo, seed):
    """Randomised parity of the five metrics: image sizes from a few pixels to non-square hundreds, 1..6 moving
    images (one constant, one equal to the fixed image), masks of random density, Gaussian widths 0..9, patch
    radii up to the image size limit, strides 1..3."""
    rng = np.random.default_rng(5000 + seed)
    kind = ["ncc", "grad-ncc", "patch-ncc", "patch-grad-ncc", "ssd"][seed % 5]
    rows, cols = int(rng.integers(7, 90)), int(rng.integers(7, 140))
    n = int(rng.integers(1, 7))
    base = rng.standard_normal((rows, cols))
    for ax in (0, 1):   # mild smoothing so that gradients are not pure noise
        base = (base + np.roll(base, 1, ax) + np.roll(base, -1, ax)) / 3.0
    fixed = (base * 4 + 6).astype(f32)
    mov = np.stack([(rng.uniform(0.2, 1.0) * fixed + rng.uniform(0.0, 1.0) * rng.standard_normal((rows, cols)) * 2).astype(f32)
                    for _ in range(n)])
    if n >= 2:
        mov[-1] = 3.25          # constant image: sigma clamp
    if n >= 3:
        mov[-2] = fixed
    mask = None
    if rng.random() < 0.5:
        mask = (rng.random((rows, cols)) < rng.uniform(0.3, 0.95)).astype(np.uint8)
        mask[rows // 2, cols // 2] = 1
    width = int(rng.choice([0, 3, 5, 7, 9]))
    rmax = max(1, min(rows, cols) // 2 - 1)
    # radius >= 2: with 3x3 patches of a heavily smoothed gradient image the reference's f32 two-pass patch
    # statistics carry ~1e-5 of rounding noise themselves (seen: 1.3e-5 at radius 1, Gaussian width 9)
    radius = int(rng.integers(2, max(min(rmax, 14), 2) + 1))
    stride = int(rng.integers(1, 4))
    cls = {"ncc": xreg_b200.ImgSimMetric2DNCCCUDA, "grad-ncc": xreg_b200.ImgSimMetric2DGradNCCCUDA,
           "patch-ncc": xreg_b200.ImgSimMetric2DPatchNCCCUDA, "patch-grad-ncc": xreg_b200.ImgSimMetric2DPatchGradNCCCUDA,
           "ssd": xreg_b200.ImgSimMetric2DSSDCUDA}[kind]
    sm = cls(ctx)
    if "grad" in kind:
        sm.set_smooth_img_before_sobel_kernel_radius(width)
    opts = xo.patch_opts(radius=radius, stride=stride)
    if "patch" in kind:
        sm.set_patch_radius(radius)
        sm.set_patch_stride(stride)
    sm.set_num_moving_images(n)
    sm.set_fixed_image(fixed)
    sm.set_mov_imgs_host_buf(np.ascontiguousarray(mov))
    if mask is not None:
        sm.set_mask(mask)
    sm.allocate_resources()
    sm.compute()
    got = sm.sim_vals()[:n].copy()
    w = xo.patch_weights(rows, cols, opts, mask=mask) if (mask is not None and "patch" in kind) else None
    if kind == "ncc":
        ref = xo.ncc(fixed, mov, mask=mask)
    elif kind == "grad-ncc":
        ref = xo.grad_ncc(fixed, mov, mask=mask, gauss_width=width)
    elif kind == "patch-ncc":
        ref = xo.patch_ncc(fixed, mov, opts, mask=mask, weights=w)
    elif kind == "patch-grad-ncc":
        ref = xo.patch_grad_ncc(fixed, mov, opts, mask=mask, weights=w, gauss_width=width)
    else:
        ref = xo.ssd(fixed, mov, mask)
    sm.close()
    assert np.all(np.isfinite(got))
    if kind == "ssd":
        assert np.all(np.abs(got - ref) <= 1e-5 * ref + 1e-9 * max(float(ref.max()), 1e-30))
    else:
        assert np.max(np.abs(got - ref)) <= 1.0e-5, (kind, rows, cols, radius, stride, width, mask is not None)


def _hdr_volume(rng, dims, kind):
    """High-dynamic-range volumes: metal next to air, where the difference-form records
    (drr.cu, pax_lerp: coefficients dA, dB, dAB rounded once; per-sample error ~ ulp of the largest corner) are weakest."""
    nx, ny, nz = dims
    if kind == 0:      # soft tissue 0.02 with metal blocks (3.0) and air pockets (0) with sharp faces
        data = np.full((nz, ny, nx), 0.02, f32)
        for _ in range(6):
            c = [int(rng.integers(0, d)) for d in (nz, ny, nx)]
            r = [int(rng.integers(1, 5)) for _ in range(3)]
            val = f32(3.0) if rng.random() < 0.5 else f32(0.0)
            data[max(0, c[0] - r[0]):c[0] + r[0], max(0, c[1] - r[1]):c[1] + r[1], max(0, c[2] - r[2]):c[2] + r[2]] = val
    elif kind == 1:    # isolated metal voxels and wires in air
        data = np.zeros((nz, ny, nx), f32)
        for _ in range(12):
            data[int(rng.integers(0, nz)), int(rng.integers(0, ny)), int(rng.integers(0, nx))] = f32(rng.choice([3.0, 10.0, 0.5]))
        data[nz // 2, ny // 2, :] = f32(7.5)      # a wire along x
        data[:, ny // 3, nx // 3] = f32(2.25)     # and one along z
    elif kind == 2:    # a CT in Hounsfield units with a metal implant, converted like the reference (HUToLinAtt)
        hu = rng.uniform(-1000.0, 1500.0, (nz, ny, nx)).astype(f32)
        hu[nz // 4:nz // 2, ny // 4:ny // 2, nx // 4:nx // 2] = f32(30000.0)      # saturated metal
        hu[:2] = f32(-1000.0)
        return hu, True
    else:              # eight decades: 1e-6 background, 1e2 inserts, checkerboard of the two on one face
        data = np.full((nz, ny, nx), 1.0e-6, f32)
        data[::2, ::2, ::2] = f32(1.0e2)
        data[nz // 2:] = f32(1.0e-6)
    return data, False


@pytest.mark.parametrize("seed", range(16 * _SCALE))
def test_high_dynamic_range_volume_matches_oracle(ctx, xo, seed):
    rng = np.random.default_rng(31000 + seed)
    dims = [int(rng.integers(6, 36)) for _ in range(3)]
    data, is_hu = _hdr_volume(rng, dims, seed % 4)
    spacing = tuple(float(s) for s in rng.uniform(0.5, 1.5, 3))
    vol = Volume(data, spacing=spacing, origin=(-10.0, 5.0, 3.0), direction=np.eye(3))
    rows, cols = int(rng.integers(24, 64)), int(rng.integers(24, 64))
    cam = CameraModel().setup(420.0, rows, cols, 1.3, 1.3)
    base = synth.nominal_pose(vol, src_to_iso=230.0, view_rot_deg=float(rng.choice([0.0, 35.0, 90.0, 140.0])))
    poses = synth.pose_population(vol, base, 4, seed=int(seed), sigma=(8, 8, 8, 3, 3, 6))
    xcam = [xo.cam_struct(cam)]
    step = float(rng.choice([0.3, 1.0, 1.0]))
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    if is_hu:
        rc.set_volumes_hu([vol], hu_lower=-1000.0)
        lin = Volume(xo.hu_to_lin_att(vol.data, -1000.0), spacing=spacing, origin=vol.origin, direction=vol.direction)
    else:
        rc.set_volume(vol)
        lin = vol
    ref, mask, steps, S = xo.drr(lin.data, lin.idx_to_phys(), xcam, to12(poses), step_size=step, want_info=True)
    rc.set_camera_model(cam)
    rc.set_ray_step_size(step)
    rc.set_num_projs(4)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    got = rc.raw_host_pixel_buf().copy()
    gmask, gsteps, gS = rc.ray_info()
    rc.close()
    np.testing.assert_array_equal(gmask, mask)
    np.testing.assert_array_equal(gsteps, steps)
    assert np.all(got[mask == 0] == ref[mask == 0])
    scale = float(ref.max())
    assert scale > 0
    # north_star: per-pixel relative error <= 1e-4.  Every pixel with a line integral above 1e-5 of the brightest one is
    # judged relatively (a ray that only grazes one metal voxel still qualifies); below that, absolutely
    sel = (mask == 1) & (ref > 1.0e-5 * scale)
    rel = np.abs(got[sel] - ref[sel]) / ref[sel]
    assert rel.max() <= 1.0e-4, (seed, float(rel.max()))
    assert np.abs(got - ref).max() <= 1.0e-6 * scale
