"""CUDA path against the REFERENCE'S OWN CODE, end to end: DRRs from the reference's ComputeLineInts<Kernel>
(xregRayCastLineIntCPU.cpp:40-292) and similarity values from its ImgSimMetric2D{NCC,GradNCC,PatchNCC,PatchGradNCC}CPU
classes, compiled from /root/reference over stand-in types into oracle/_ref/ (oracle/ref_pin/build_ref_slice.py), with the
real OpenCV (cv2) behind cv::GaussianBlur / cv::Sobel.  This is the closest thing to "the reference run here" that this
image allows; the tolerances are north_star's: clip masks exact, DRR <= 1e-4 relative, similarity <= 1e-5 absolute."""
import numpy as np
import pytest

from oracle.ref_pin import ref_slice
from xreg_b200 import regi, synth
from xreg_b200.geometry import to12

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_slice.available(), reason="neither the reference checkout nor a built oracle/_ref slice")]
f32 = np.float32


def _cv2_filters():
    cv2 = pytest.importorskip("cv2")
    return (lambda img, k: cv2.GaussianBlur(img, (k, k), 0)), (lambda img, dx, dy: cv2.Sobel(img, -1, dx, dy))


@pytest.mark.parametrize("metric", ["patch-grad-ncc", "grad-ncc", "ncc", "patch-ncc"])
def test_cuda_path_matches_the_reference_code(ctx, xo, small_scene, metric):
    vol, cam, nominal = small_scene
    pop = synth.pose_population(vol, nominal, 6)
    xcam = [xo.cam_struct(cam)]
    # reference code: DRRs
    ref_drr = ref_slice.compute_line_ints(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcam, to12(pop))
    fixed = synth.add_noise(ref_drr[0])
    mask = synth.circular_mask(cam.num_det_rows, cam.num_det_cols) if metric in ("patch-grad-ncc", "ncc") else None
    # CUDA path
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric=metric, max_pop=6, patch_radius=6,
                                 masks=[mask] if mask is not None else None)
    sims = fn(pop)
    got = fn.rc.raw_host_pixel_buf().copy()
    fn.close()
    # DRR parity against the reference code
    assert np.array_equal(got == 0, ref_drr == 0)          # rays that miss / cross only air: exactly zero in both
    sel = ref_drr > 1e-3 * ref_drr.max()
    assert (np.abs(got[sel] - ref_drr[sel]) / ref_drr[sel]).max() <= 1e-4
    # metric parity against the reference classes, fed with the reference's DRRs, real OpenCV inside
    ref_slice.set_filters(*_cv2_filters())
    opts = xo.patch_opts(radius=6)
    if metric == "patch-grad-ncc":
        ref = ref_slice.patch_grad_ncc(fixed, ref_drr, opts, mask=mask, gauss_width=5)
    elif metric == "grad-ncc":
        ref = ref_slice.grad_ncc(fixed, ref_drr, mask=mask, gauss_width=5)
    elif metric == "ncc":
        ref = ref_slice.ncc(fixed, ref_drr, mask=mask)
    else:
        ref = ref_slice.patch_ncc(fixed, ref_drr, opts, mask=mask)[0]
    assert np.max(np.abs(sims - ref)) <= 1e-5, (metric, sims, ref)
    assert int(np.argmin(sims)) == 0 == int(np.argmin(ref))


def test_config_c1_full_size_against_the_reference_code(ctx, xo):
    """BASELINE.json configs[0] verbatim -- "single line-integral DRR of synthetic 256^3 float CT at 256x256 detector via
    RayCasterLineIntCPU + NCC" -- with the reference's own ComputeLineInts and NCC class code as the CPU side."""
    vol = synth.make_volume(256, 256, 256)
    cam = synth.make_camera(256)
    nominal = synth.nominal_pose(vol)
    pop = synth.pose_population(vol, nominal, 2)
    xcam = [xo.cam_struct(cam)]
    ref_drr = ref_slice.compute_line_ints(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcam, to12(pop))
    fixed = synth.add_noise(ref_drr[0])
    fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric="ncc", max_pop=2)
    sims = fn(pop)
    got = fn.rc.raw_host_pixel_buf().copy()
    gmask, gsteps, gS = fn.rc.ray_info()
    fn.close()
    assert ref_drr.max() > 1.0 and (ref_drr > 0).mean() > 0.3
    assert np.array_equal(got == 0, ref_drr == 0)
    sel = ref_drr > 1e-3 * ref_drr.max()
    assert (np.abs(got[sel] - ref_drr[sel]) / ref_drr[sel]).max() <= 1e-4
    ref = ref_slice.ncc(fixed, ref_drr)
    assert np.max(np.abs(sims - ref)) <= 1e-5, (sims, ref)
    # and the oracle is that code, bit for bit, at this size too
    o = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(pop))
    assert o.tobytes() == ref_drr.tobytes()


def test_objective_from_se3_parameters_against_the_reference_exp_map(ctx, xo, small_scene):
    """xrc_obj_fn_se3 composes pose_p = pre * ExpSE3(x_p) * post inside the library; here the same
    poses are composed with the REFERENCE's own ExpSE3 / ExpSO3 lines (lib/transforms/xregRigidUtils.cpp:40-85,
    xregRotUtils.cpp:33-105, oracle/_ref/libxreg_refslice_se3.so), ray cast and scored by the oracle, and the library's
    one-call objective must agree within the north_star similarity tolerance (the two exp maps agree to f32 rounding,
    <= 2e-6 per matrix entry: tests/test_oracle_ref_slice.py)."""
    from xreg_b200 import regi
    from xreg_b200.geometry import exp_se3

    vol, cam, nominal = small_scene
    c = np.asarray(vol.origin) + 0.5 * (np.asarray(vol.dims) - 1.0) * np.asarray(vol.spacing)
    pre, post = np.eye(4, dtype=f32), np.eye(4, dtype=f32)
    pre[:3, 3] = c
    post[:3, 3] = -c
    post = (post @ nominal).astype(f32)
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.normal(0, 0.08, (8, 3)), rng.normal(0, 5.0, (8, 3))], axis=1).astype(f32)
    x[0] = 0
    ref_poses = []
    for xi in x:
        E = np.vstack([ref_slice.exp_se3(xi).reshape(3, 4), [0, 0, 0, 1]]).astype(f32)
        ref_poses.append((pre.astype(np.float64) @ E.astype(np.float64) @ post.astype(np.float64)).astype(f32))
    ref_poses = np.stack(ref_poses)
    xcam = [xo.cam_struct(cam)]
    drr = xo.drr(vol.data, vol.idx_to_phys(), xcam, to12(ref_poses))
    fixed = synth.add_noise(drr[0])
    for metric, ofn in (("grad-ncc", lambda d: xo.grad_ncc(fixed, d)), ("ncc", lambda d: xo.ncc(fixed, d)),
                        ("patch-grad-ncc", lambda d: xo.patch_grad_ncc(fixed, d, xo.patch_opts(radius=6)))):
        fn = regi.Intensity2D3DObjFn(ctx, vol, [cam], [fixed], metric=metric, max_pop=8, patch_radius=6)
        got = fn.eval_se3(x, pre, post)
        assert np.max(np.abs(got - ofn(drr))) <= 1e-5, metric
        assert int(np.argmin(got)) == 0
        fn.close()
