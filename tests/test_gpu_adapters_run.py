"""the xReg adapter classes LINKED and RUN.  oracle/_ref/xreg_adapter_driver (tests/xreg_link/) is
adapters/xreg/*.cpp linked with the reference's own RayCaster / ImgSimMetric2D / PatchCommon / CombineMean / CameraModel
sources; it drives xreg::RayCasterLineIntCUDA and xreg::ImgSimMetric2D*CUDA behind the reference's base classes in the
order Intensity2D3DRegi::setup() / obj_fn() use them.  Its projections and similarity values must equal the Python host
mirror's bit for bit (same library underneath, so any difference is an adapter bug) and the CPU oracle's within the
north_star tolerances."""
import os
import struct
import subprocess

import numpy as np
import pytest

import xreg_b200
from xreg_b200 import regi, synth
from xreg_b200.geometry import CameraModel, to12

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "xreg_adapter_driver")
KINDS = {"ncc": 0, "grad-ncc": 1, "patch-ncc": 2, "patch-grad-ncc": 3, "ssd": 4}
f32 = np.float32


def _via_intrins(cam):
    """The same camera set up from its intrinsic / extrinsic matrices, the overload the driver uses
    (CameraModel::setup(intrins, extrins, ...), xregPerspectiveXform.cpp:302-334: focal length re-derived from K)."""
    c = CameraModel(coord_frame_type=cam.coord_frame_type)
    c.setup_intrins_extrins(cam.intrins, np.asarray(cam.extrins, f32), cam.num_det_rows, cam.num_det_cols,
                            cam.det_row_spacing, cam.det_col_spacing)
    return c


def _driver():
    if os.path.isdir("/root/reference"):
        from tests.xreg_link import build_link

        build_link.build()
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/xreg_adapter_driver not built (needs /root/reference at build time)")
    return DRIVER


def write_input(path, vols, cams, fixed, masks, poses, kind, patch_radius=5, patch_stride=1, gauss=5, step=1.0,
                static_pose=None, weight_flags=2 | 4, subset=()):
    """poses: (n_evals, n_moving, pop, 4, 4).  vols: moving volumes, then the static one when static_pose is given."""
    n_evals, n_moving, pop = poses.shape[:3]
    v0 = vols[0]
    rows, cols = cams[0].num_det_rows, cams[0].num_det_cols
    with open(path, "wb") as f:
        f.write(b"XRLK")
        f.write(struct.pack("<3I", len(cams), pop, n_evals))
        f.write(struct.pack("<3I", *v0.dims))
        f.write(struct.pack("<2I", rows, cols))
        f.write(struct.pack("<9I", KINDS[kind], patch_radius, patch_stride, gauss, 1 if masks else 0, n_moving,
                            1 if static_pose is not None else 0, weight_flags, len(subset)))
        f.write(struct.pack("<f", step))
        f.write(np.asarray(v0.spacing, np.float64).tobytes())
        f.write(np.asarray(v0.origin, np.float64).tobytes())
        f.write(np.asarray(v0.direction, np.float64).reshape(9).tobytes())
        for v in vols:
            f.write(np.ascontiguousarray(v.data, f32).tobytes())
        for c in cams:
            f.write(np.ascontiguousarray(c.intrins, f32).tobytes())
            f.write(np.ascontiguousarray(c.extrins, f32).reshape(16).tobytes())
            f.write(struct.pack("<2fI", c.det_row_spacing, c.det_col_spacing, c.coord_frame_type))
        for im in fixed:
            f.write(np.ascontiguousarray(im, f32).tobytes())
        for m in masks or []:
            f.write(np.ascontiguousarray(m, np.uint8).tobytes())
        f.write(np.asarray(subset, np.uint64).tobytes())
        if static_pose is not None:
            f.write(np.ascontiguousarray(static_pose, f32).reshape(16).tobytes())
        f.write(np.ascontiguousarray(poses, f32).tobytes())


def read_output(path, n_evals, n_views, pop, rows, cols):
    raw = np.fromfile(path, dtype=np.uint8)
    off = 0

    def take(n, dt=f32):
        nonlocal off
        a = raw[off:off + n * np.dtype(dt).itemsize].view(dt)
        off += n * np.dtype(dt).itemsize
        return a.copy()

    sims, per_view = [], []
    for _ in range(n_evals):
        sims.append(take(pop))
        per_view.append(take(n_views * pop).reshape(n_views, pop))
    projs = take(n_views * pop * rows * cols).reshape(n_views * pop, rows, cols)
    last = take(rows * cols).reshape(rows, cols)
    spacing = take(2)
    first_ocv = take(rows * cols).reshape(rows, cols)
    n_patches = int(take(1, np.uint64)[0])
    assert off == raw.size
    return np.stack(sims), np.stack(per_view), projs, last, spacing, first_ocv, n_patches


def run_driver(tmp_path, **kw):
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_input(inp, **kw)
    r = subprocess.run([_driver(), inp, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    poses, cams = kw["poses"], kw["cams"]
    return read_output(out, poses.shape[0], len(cams), poses.shape[2], cams[0].num_det_rows, cams[0].num_det_cols)


def test_driver_fails_loudly_without_a_gpu(tmp_path):
    """(runs everywhere) the binary exists where the reference was available at build time, parses its input with the
    reference's own CameraModel::setup, and -- on a box without a GPU -- stops at xrc_ctx_create with the library's
    message instead of falling back to anything."""
    import torch

    vol = synth.make_volume(16, 16, 12)
    cam = CameraModel().setup(300.0, 12, 16, 4.0, 4.0)
    poses = synth.pose_population(vol, synth.nominal_pose(vol, src_to_iso=150.0), 2)[None, None]
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_input(inp, vols=[vol], cams=[cam], fixed=[np.ones((12, 16), f32)], masks=None, poses=poses, kind="ncc")
    r = subprocess.run([_driver(), inp, out], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ncc", "grad-ncc", "patch-ncc", "patch-grad-ncc", "ssd"])
def test_adapters_behind_the_reference_base_classes(ctx, xo, small_scene, tmp_path, kind):
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    cams = [_via_intrins(cam), _via_intrins(cam2)]
    rows, cols = cam.num_det_rows, cam.num_det_cols
    pops = np.stack([synth.pose_population(vol, nominal, 5, seed=s) for s in (3, 4, 5)])   # three evaluations
    xcams = [xo.cam_struct(c) for c in cams]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), [xc], to12(pops[0][:1]))[0], seed=s) for s, xc in enumerate(xcams)]
    mask = synth.circular_mask(rows, cols, 0.85)
    for masks in (None, [mask, mask]):
        sims, per_view, projs, last, spacing, first_ocv, n_patches = run_driver(
            tmp_path, vols=[vol], cams=cams, fixed=fixed, masks=masks, poses=pops[:, None], kind=kind, patch_radius=6)
        # the Python host mirror over the same library: bit for bit
        fn = regi.Intensity2D3DObjFn(ctx, vol, cams, fixed, metric=kind, max_pop=5, patch_radius=6,
                                     masks=masks if masks else None)
        for e in range(3):
            got = fn(pops[e])
            np.testing.assert_array_equal(sims[e], got)
            np.testing.assert_array_equal(per_view[e], np.stack([sm.sim_vals()[:5] for sm in fn.sims]))
        mirror_projs = fn.rc.raw_host_pixel_buf()
        np.testing.assert_array_equal(projs, mirror_projs)
        np.testing.assert_array_equal(last, mirror_projs[-1])
        np.testing.assert_array_equal(first_ocv, mirror_projs[0])
        # proj(last) is a view of the LAST camera's image with that camera's spacing (xregRayCastBaseCPU.cpp:90-118)
        assert spacing[0] == f32(cams[-1].det_col_spacing) and spacing[1] == f32(cams[-1].det_row_spacing)
        if kind.startswith("patch"):
            assert n_patches == (rows - 12) * (cols - 12)
        fn.close()
        # the CPU oracle: north_star tolerances
        p12, ci = xo.distribute_xforms(to12(pops[2]), 2)
        ref = xo.drr(vol.data, vol.idx_to_phys(), xcams, p12, cam_idx=ci)
        sel = ref > 1e-3 * ref.max()
        assert np.array_equal(projs == 0, ref == 0)
        assert np.max(np.abs(projs[sel] - ref[sel]) / ref[sel]) <= 1e-4
        mk = masks[0] if masks else None
        ofn = {"ncc": lambda f, d: xo.ncc(f, d, mask=mk), "grad-ncc": lambda f, d: xo.grad_ncc(f, d, mask=mk),
               "ssd": lambda f, d: xo.ssd(f, d, mask=mk),
               "patch-ncc": lambda f, d: xo.patch_ncc(f, d, xo.patch_opts(radius=6), mask=mk,
                                                      weights=xo.patch_weights(rows, cols, xo.patch_opts(radius=6), mask=mk)),
               "patch-grad-ncc": lambda f, d: xo.patch_grad_ncc(f, d, xo.patch_opts(radius=6), mask=mk,
                                                                weights=xo.patch_weights(rows, cols, xo.patch_opts(radius=6), mask=mk))}[kind]
        oref = xo.combine_mean(np.stack([ofn(fixed[0], ref[:5]), ofn(fixed[1], ref[5:])]))
        tol = 1e-5 if kind != "ssd" else 1e-5 * max(1.0, float(np.max(np.abs(oref))))
        assert np.max(np.abs(sims[2] - oref)) <= tol, kind


@pytest.mark.gpu
def test_adapter_static_volume_background_survives_repeated_evaluations(ctx, xo, small_scene, tmp_path):
    """The ADVICE r1 case: Intensity2D3DRegi::obj_fn toggles set_use_bg_projs(true) -> compute(vol 0) -> (false) ->
    compute(vol 1) on EVERY evaluation (xregIntensity2D3DRegi.cpp:594-629).  The static volume's background must be
    in every evaluation's projections, not only the first one's."""
    vol, cam, nominal = small_scene
    bone = xreg_b200.Volume(np.where(vol.data >= 0.045, vol.data, 0.0).astype(f32), vol.spacing, vol.origin, vol.direction)
    soft = xreg_b200.Volume(np.where(vol.data < 0.045, vol.data, 0.0).astype(f32), vol.spacing, vol.origin, vol.direction)
    other = xreg_b200.Volume((0.5 * vol.data[::-1]).copy().astype(f32), vol.spacing, vol.origin, vol.direction)
    pops = np.stack([np.stack([synth.pose_population(vol, nominal, 4, seed=10 * e + o) for o in range(2)]) for e in range(3)])
    cam = _via_intrins(cam)
    xcam = [xo.cam_struct(cam)]
    static_pose = nominal
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pops[0, 0, :1]))[0])]
    sims, per_view, projs, *_ = run_driver(tmp_path, vols=[bone, other, soft], cams=[cam], fixed=fixed, masks=None,
                                           poses=pops, kind="grad-ncc", static_pose=static_pose)
    bg = xo.drr(soft.data, soft.idx_to_phys(), xcam, to12(static_pose[None]))[0]
    for e in range(3):
        buf = np.repeat(bg[None], 4, axis=0).copy()
        xo.drr(bone.data, bone.idx_to_phys(), xcam, to12(pops[e, 0]), buf=buf)
        xo.drr(other.data, other.idx_to_phys(), xcam, to12(pops[e, 1]), buf=buf)
        ref = xo.grad_ncc(fixed[0], buf)
        assert np.max(np.abs(sims[e] - ref)) <= 1e-5, "evaluation %d lost the static background" % e
    sel = buf > 1e-3 * buf.max()
    assert np.max(np.abs(projs[sel] - buf[sel]) / buf[sel]) <= 1e-4


@pytest.mark.gpu
def test_adapter_patch_subset_through_the_reference_mixin(ctx, xo, small_scene, tmp_path):
    """set_patches_to_use through the reference's own ImgSimMetric2DPatchCommon (reached by dynamic_cast, as the apps do):
    the adapter hands the local patch list to the library before every compute()."""
    vol, cam, nominal = small_scene
    cam = _via_intrins(cam)
    rows, cols = cam.num_det_rows, cam.num_det_cols
    pops = synth.pose_population(vol, nominal, 4, seed=8)[None, None]
    xcam = [xo.cam_struct(cam)]
    fixed = [synth.add_noise(xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pops[0, 0, :1]))[0])]
    n_p = xo.num_patches(rows, cols, 6, 1)
    sub = np.random.default_rng(3).integers(0, n_p, size=200)
    sims, per_view, projs, _, _, _, n_patches = run_driver(tmp_path, vols=[vol], cams=[cam], fixed=fixed, masks=None, poses=pops,
                                                          kind="patch-grad-ncc", patch_radius=6, subset=sub)
    assert n_patches == 200          # ImgSimMetric2DPatchCommon::num_patches() of a pinned list
    ref = xo.patch_ncc_subset(fixed[0], xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pops[0, 0])), xo.patch_opts(radius=6), sub,
                              gauss_width=5)
    assert np.max(np.abs(sims[0] - ref)) <= 1e-5
