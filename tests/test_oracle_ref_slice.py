"""Pins the oracle to the reference's OWN code for the geometric core of the DRR path.

oracle/ref_pin/build_ref_slice.py compiles, from /root/reference where it lies, the reference's RayRectIntersect
(lib/spatial/xregSpatialPrimitives.cpp:175-222), CameraModel::ind_pt_to_phys_det_pt
(lib/transforms/xregPerspectiveXform.cpp:391-414) and the line-integral kernels + ComputeLineInts<Kernel>
(lib/ray_cast/xregRayCastLineIntCPU.cpp:40-292) over functional stand-ins for the Eigen / ITK / TBB types
(oracle/ref_pin/ref_pin_prelude.h: the un-vendored dependencies restated with the conventions DESIGN.md lists).
The oracle's C restatement (oracle/xreg_oracle.c: xo_drr) must agree with that code BIT FOR BIT: every pixel of every
projection, both kernels, all camera frame types, oblique volumes, tiny volumes, cameras inside / missing the volume,
step sizes, REPLACE and ACCUM stores.  A transcription error in the restatement (operation order, a cast, a comparison,
the nudge, the step count, the loop bounds) shows up here.

Runs where the reference checkout exists (it rebuilds the slice) or where the built oracle/_ref/libxreg_refslice.so
was shipped with the snapshot; skipped otherwise."""
import numpy as np
import pytest

from oracle.ref_pin import ref_slice
from xreg_b200 import synth
from xreg_b200.geometry import CameraModel, to12

from .test_gpu_fuzz import _scene

pytestmark = pytest.mark.skipif(not ref_slice.available(), reason="neither the reference checkout nor a built oracle/_ref slice")
f32 = np.float32


def test_ray_rect_intersect_is_the_reference_code(xo):
    """The slab test alone: random segments against random boxes, axis-parallel and grazing cases included, compared
    through one-pixel DRR set-ups is indirect -- here the reference function is called directly and the oracle's
    restatement is exercised through xo.drr's clip mask below; this test fixes the function's known answers."""
    hit, t0, t1 = ref_slice.ray_rect_intersect([0, 0, 0], [9, 9, 9], [-5, 4, 4], [20, 0, 0])
    assert hit and t0 == f32(0.25) and t1 == f32(0.7)
    hit, _, _ = ref_slice.ray_rect_intersect([0, 0, 0], [9, 9, 9], [-5, 10, 4], [20, 0, 0])       # parallel, outside a slab
    assert not hit
    hit, t0, t1 = ref_slice.ray_rect_intersect([0, 0, 0], [9, 9, 9], [4, 4, 4], [1, 1, 1])        # starts inside
    assert hit and t0 == 0 and t1 == 1
    hit, _, _ = ref_slice.ray_rect_intersect([0, 0, 0], [9, 9, 9], [-5, 4, 4], [2, 0, 0])         # segment ends before the box
    assert not hit
    hit, t0, t1 = ref_slice.ray_rect_intersect([0, 0, 0], [9, 9, 9], [-5, 4, 4], [2, 0, 0], limit_to_segment=False)
    assert hit and t0 == f32(2.5) and t1 == f32(7.0)


@pytest.mark.parametrize("frame", [0, 1, 2])
def test_detector_points_are_the_reference_code(xo, frame):
    """ind_pt_to_phys_det_pt of the reference == the detector points the oracle's DRR uses: checked through a camera with
    non-trivial extrinsics for every frame type (the oracle's camera set-up fills intrins_inv / extrins_inv / pinhole)."""
    rng = np.random.default_rng(frame)
    E = np.eye(4)
    E[:3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    E[:3, 3] = rng.uniform(-40, 40, 3)
    K = np.array([[-1200.0, 0.3, 35.5], [0, -1180.0, 28.25], [0, 0, 1]])
    cam = xo.cam_setup(K, E, 64, 72, 0.8, 0.9, frame_type=frame)
    # the reference maps (col, row) -> camera-world point; linear in (col, row) up to f32 rounding: spot values + consistency
    p00 = ref_slice.ind_pt_to_phys_det_pt(cam, 0, 0)
    p10 = ref_slice.ind_pt_to_phys_det_pt(cam, 1, 0)
    p01 = ref_slice.ind_pt_to_phys_det_pt(cam, 0, 1)
    p = ref_slice.ind_pt_to_phys_det_pt(cam, 37, 21)
    assert np.allclose(p, p00 + 37 * (p10 - p00) + 21 * (p01 - p00), rtol=0, atol=2e-3)
    # a one-voxel-thick slab orthogonal to nothing in particular: the DRR comparison below is the bitwise statement


def _both(xo, vol, cam, poses, step, kernel_id, cam_idx=None, buf=None, interp=0):
    xcams = cam if isinstance(cam, list) else [xo.cam_struct(cam)]
    p12 = to12(poses)
    a = xo.drr(vol.data, vol.idx_to_phys(), xcams, p12, cam_idx=cam_idx, step_size=step, kernel_id=kernel_id,
               buf=None if buf is None else buf.copy(), n_threads=1, interp=interp)
    b = ref_slice.compute_line_ints(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcams, p12, cam_idx=cam_idx,
                                    step_size=step, kernel_id=kernel_id, buf=None if buf is None else buf.copy(), interp=interp)
    return a, b


@pytest.mark.parametrize("seed", range(0, 48, 3))
def test_oracle_nearest_neighbour_drr_equals_the_reference_code(xo, seed):
    """kRAY_CAST_INTERP_NN (xregRayCastLineIntCPU.cpp:128-130): the reference's ComputeLineInts with the
    nearest-neighbour interpolator (ITK's rounding stated as floor(x + 0.5)) against xo_drr_interp, bit for bit; and
    the two interpolators really differ."""
    vol, cam, poses, step, kernel_id, kind = _scene(seed)
    a, b = _both(xo, vol, cam, poses, step, kernel_id, interp=1)
    assert a.tobytes() == b.tobytes()
    lin, _ = _both(xo, vol, cam, poses, step, kernel_id, interp=0)
    if vol.data.size > 8 and np.ptp(vol.data) > 0 and a.max() > 0:
        assert a.tobytes() != lin.tobytes()


@pytest.mark.parametrize("seed", range(48))
def test_oracle_drr_equals_the_reference_code_on_random_scenes(xo, seed):
    """The 48 scenes of the GPU fuzz test (tests/test_gpu_fuzz.py::_scene)."""
    vol, cam, poses, step, kernel_id, kind = _scene(seed)
    a, b = _both(xo, vol, cam, poses, step, kernel_id)
    assert a.tobytes() == b.tobytes()
    # ACCUM on top of a previous projection (the store `buf = K(buf, val)`)
    prev = (np.abs(a) * f32(0.5) + f32(0.125)).astype(f32)
    a2, b2 = _both(xo, vol, cam, poses, step, kernel_id, buf=prev)
    assert a2.tobytes() == b2.tobytes()


@pytest.mark.parametrize("seed", range(0, 48, 2))
def test_oracle_depth_equals_the_reference_code(xo, seed):
    """RayCasterDepthCPU (xregRayCastDepthCPU.cpp:42-229): the reference's RayCastDepthFn -- first sample >= threshold
    along the unlimited ray, step-halving refinement, distance from the pinhole through the inverse camera -> index
    transform, min store -- against xo_depth, bit for bit, linear and nearest-neighbour interpolation, several
    thresholds and refinement counts, on top of kRAY_CAST_MAX_DEPTH and on top of a previous depth image."""
    vol, cam, poses, step, kernel_id, kind = _scene(seed)
    xcams = [xo.cam_struct(cam)]
    p12 = to12(poses)
    vmax = float(vol.data.max())
    hits = 0
    for interp in (0, 1):
        for frac, nb in ((0.3, 0), (0.6, 4), (0.05, 20), (2.0, 3)):
            thr = frac * vmax if vmax > 0 else 0.5
            a = xo.depth(vol.data, vol.idx_to_phys(), xcams, p12, step_size=step, interp=interp, thresh=thr, n_backtrack=nb,
                         n_threads=1)
            b = ref_slice.compute_depth(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcams, p12, step_size=step,
                                        interp=interp, thresh=thr, n_backtrack=nb)
            assert a.tobytes() == b.tobytes()
            hits += int(np.count_nonzero(a < 1.0e36))
            if frac >= 2.0:
                assert np.all(a == xo.RAY_CAST_MAX_DEPTH)      # nothing reaches a threshold above the maximum
            prev = np.where(a < 1.0e36, a * f32(0.5), f32(40.0)).astype(f32)
            a2 = xo.depth(vol.data, vol.idx_to_phys(), xcams, p12, step_size=step, interp=interp, thresh=thr, n_backtrack=nb,
                          buf=prev.copy(), n_threads=1)
            b2 = ref_slice.compute_depth(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcams, p12, step_size=step,
                                         interp=interp, thresh=thr, n_backtrack=nb, buf=prev.copy())
            assert a2.tobytes() == b2.tobytes() and np.all(a2 <= prev)
    if vmax > 0 and seed in (0, 2, 4, 6):
        assert hits > 0        # the comparison is not vacuous: surfaces are found


def test_oracle_drr_equals_the_reference_code_on_the_test_scene(xo, small_scene):
    """The 64x64x48 anisotropic phantom of the GPU parity tests, two cameras, camera-major association."""
    vol, cam, nominal = small_scene
    cam2 = CameraModel().setup(380.0, cam.num_det_rows, cam.num_det_cols, 1.7, 1.4)
    pop = synth.pose_population(vol, nominal, 3)
    xcams = [xo.cam_struct(cam), xo.cam_struct(cam2)]
    poses = np.concatenate([pop, pop])
    cam_idx = np.array([0, 0, 0, 1, 1, 1], np.uint32)
    for step in (1.0, 0.37):
        a, b = _both(xo, vol, xcams, poses, step, 0, cam_idx=cam_idx)
        assert a.max() > 0 and a.tobytes() == b.tobytes()
    # the comparison is not vacuous: one ulp on the step size changes the bytes
    c = ref_slice.compute_line_ints(vol.data, xo.affine_inverse(vol.idx_to_phys()), xcams, to12(poses), cam_idx=cam_idx,
                                    step_size=float(np.nextafter(f32(0.37), f32(1.0))))
    assert c.tobytes() != b.tobytes()


# ---- the patch-NCC metric: the reference's own class code ----------------------------------------------------------------
def _metric_case(seed):
    rng = np.random.default_rng(9000 + seed)
    rows, cols = int(rng.integers(7, 70)), int(rng.integers(7, 90))
    n = int(rng.integers(1, 5))
    base = rng.standard_normal((rows, cols))
    for ax in (0, 1):
        base = (base + np.roll(base, 1, ax) + np.roll(base, -1, ax)) / 3.0
    fixed = (base * 4 + 6).astype(f32)
    mov = np.stack([(rng.uniform(0.2, 1.0) * fixed + rng.uniform(0.0, 1.0) * rng.standard_normal((rows, cols)) * 2).astype(f32)
                    for _ in range(n)])
    if n >= 2:
        mov[-1] = 3.25          # constant image: every patch takes the sigma clamp
    if n >= 3:
        mov[-2] = fixed
    mask = None
    if seed % 2 == 1:
        mask = (rng.random((rows, cols)) < rng.uniform(0.3, 0.95)).astype(np.uint8)
        mask[rows // 2, cols // 2] = 1
    rmax = max(1, min(rows, cols) // 2 - 1)
    radius = int(rng.integers(1, max(min(rmax, 12), 1) + 1))
    stride = int(rng.integers(1, 4))
    wgt_img = rng.uniform(0.0, 2.0, (rows, cols)).astype(f32) if seed % 5 == 3 else None
    return fixed, mov, mask, radius, stride, wgt_img


@pytest.mark.parametrize("seed", range(40))
def test_oracle_patch_ncc_equals_the_reference_class_code(xo, seed):
    """ImgSimMetric2DPatchNCCCPU::{allocate_resources, compute, process_mask}, ImgSimMetric2DPatchCommon::{setup_patches,
    compute_weights, patch_indices_to_use} and detail::ComputePatchMeanStdDev, compiled from the reference's own lines
    (xregImgSimMetric2DPatchNCCCPU.cpp:34-72,74-300,332-441,558-619; xregImgSimMetric2DPatchCommon.cpp:35-54,180-184,
    256-493) over cv::Mat / itk::Image stand-ins that carry no arithmetic: patch grid, patch weights, per-patch values
    and image scores of the oracle are the reference's, bit for bit, for every option combination the CUDA path accepts."""
    fixed, mov, mask, radius, stride, wgt_img = _metric_case(seed)
    rows, cols = fixed.shape
    combos = [dict(), dict(compute_mean=True), dict(weight_sims=False), dict(normalize=False)]
    if mask is not None:
        combos += [dict(mask_stats=True), dict(mask_weighting=False), dict(mask_stats=True, mask_weighting=False, normalize=False)]
    for kw in combos:
        opts = xo.patch_opts(radius=radius, stride=stride, **kw)
        w_or = xo.patch_weights(rows, cols, opts, mask=mask, wgt_img=wgt_img)
        need_w = (mask is not None and opts.use_mask_for_weighting) or wgt_img is not None
        s_or, p_or = xo.patch_ncc(fixed, mov, opts, mask=mask, weights=w_or if need_w else None, want_patch_sims=True, n_threads=1)
        s_rf, w_rf, p_rf = ref_slice.patch_ncc(fixed, mov, opts, mask=mask, wgt_img=wgt_img, want_patch_sims=True)
        assert w_or.tobytes() == w_rf.tobytes(), kw
        assert s_or.tobytes() == s_rf.tobytes(), (kw, s_or, s_rf)
        if opts.weight_patch_sims:
            live = np.abs(w_rf) > 1.0e-6      # patches the reference skips keep whatever the vector held
            assert np.array_equal(p_or[:, live], p_rf[:, live]), kw
        else:
            assert p_or.tobytes() == p_rf.tobytes(), kw


@pytest.mark.parametrize("seed", range(16))
def test_oracle_patch_subset_equals_the_reference_class_code(xo, seed):
    """ImgSimMetric2DPatchCommon::set_patches_to_use (xregImgSimMetric2DPatchCommon.cpp:231-235, cut from the
    reference) followed by the reference's ImgSimMetric2DPatchNCCCPU::compute over patch_inds_to_use_: the oracle's
    xo_patch_ncc_subset gives the same scores bit for bit -- unordered subsets, repeated indices (the random patches are
    drawn with replacement), a single patch, the whole grid in reverse."""
    fixed, mov, mask, radius, stride, wgt_img = _metric_case(seed)
    rows, cols = fixed.shape
    rng = np.random.default_rng(77 + seed)
    n_p = xo.num_patches(rows, cols, radius, stride)
    subsets = [rng.integers(0, n_p, size=max(1, n_p // 3)), np.array([int(rng.integers(0, n_p))]), np.arange(n_p)[::-1],
               np.repeat(rng.integers(0, n_p, size=3), 2)]
    combos = [dict(), dict(compute_mean=True), dict(weight_sims=False)]
    if mask is not None:
        combos += [dict(mask_stats=True)]
    for kw in combos:
        opts = xo.patch_opts(radius=radius, stride=stride, **kw)
        w_or = xo.patch_weights(rows, cols, opts, mask=mask, wgt_img=wgt_img)
        need_w = (mask is not None and opts.use_mask_for_weighting) or wgt_img is not None
        for sub in subsets:
            s_or = xo.patch_ncc_subset(fixed, mov, opts, sub, mask=mask, weights=w_or if need_w else None, n_threads=1)
            s_rf = ref_slice.patch_ncc_subset(fixed, mov, opts, sub, mask=mask, wgt_img=wgt_img)
            assert np.array_equal(s_or, s_rf, equal_nan=True), (kw, len(sub), s_or, s_rf)
    # the full grid in natural order is the plain metric
    opts = xo.patch_opts(radius=radius, stride=stride)
    w_or = xo.patch_weights(rows, cols, opts, mask=mask, wgt_img=wgt_img)
    need_w = (mask is not None and opts.use_mask_for_weighting) or wgt_img is not None
    full = xo.patch_ncc_subset(fixed, mov, opts, np.arange(n_p), mask=mask, weights=w_or if need_w else None, n_threads=1)
    assert np.array_equal(full, xo.patch_ncc(fixed, mov, opts, mask=mask, weights=w_or if need_w else None, n_threads=1), equal_nan=True)


def test_patch_mean_std_is_the_reference_code(xo):
    rng = np.random.default_rng(4)
    img = rng.standard_normal((20, 24)).astype(f32) * 3 + 1
    mask = (rng.random((20, 24)) < 0.6).astype(np.uint8)
    mean, sd, n = ref_slice.patch_mean_std(img, None, 3, 5, 7, False)
    p = img[3:10, 5:12].astype(np.float64)
    assert n == 49 and abs(mean - p.mean()) < 1e-5 and abs(sd - p.std(ddof=1)) < 1e-5
    mean, sd, n = ref_slice.patch_mean_std(img, mask, 3, 5, 7, True)
    sel = mask[3:10, 5:12] > 0
    assert n == int(sel.sum()) and abs(mean - p[sel].mean()) < 1e-5 and abs(sd - p[sel].std(ddof=1)) < 1e-5
    mean, sd, n = ref_slice.patch_mean_std(np.full((9, 9), 2.5, f32), None, 0, 0, 9, False)
    assert mean == f32(2.5) and sd == f32(1.0e-6) and n == 81          # the clamp


def test_oracle_hu_to_lin_att_equals_the_reference_filter_code(xo):
    """HUToLinAttFilter::GenerateData (lib/image/xregHUToLinAtt.cpp:45-69) over flat-image / iterator stand-ins."""
    rng = np.random.default_rng(11)
    hu = np.concatenate([rng.uniform(-1200, 3000, 50000), [-1000.0, -999.99994, -1000.0001, 0.0, 1e-30, -3e4, 6e4]]).astype(f32)
    for lower in (-1000.0, -800.0, 150.0):
        a = xo.hu_to_lin_att(hu, lower)
        b = ref_slice.hu_to_lin_att(hu, lower)
        assert a.tobytes() == b.tobytes() and a.max() > 0 and (a == 0).any()


@pytest.mark.parametrize("seed", range(30))
def test_oracle_ncc_equals_the_reference_class_code(xo, seed):
    """ImgSimMetric2DNCCCPU::{allocate_resources, compute, process_mask} and its helper templates
    (xregImgSimMetric2DNCCCPU.cpp:29-132,134-236), compiled from the reference's own lines.  With a mask the reference
    computes everything in plain scalar loops: pinned without any convention.  Without a mask it calls three Eigen
    reductions (mean, sum of squares, dot), for which the stand-in follows the documented Eigen 3.3 SSE reduction shape --
    the same convention the oracle states: that path is pinned up to it (image sizes cover every tail length)."""
    rng = np.random.default_rng(7000 + seed)
    rows, cols = int(rng.integers(1, 40)), int(rng.integers(2, 60))
    n = int(rng.integers(1, 5))
    fixed = (rng.standard_normal((rows, cols)) * 3 + 5).astype(f32)
    mov = np.stack([(rng.uniform(-1.0, 1.0) * fixed + rng.standard_normal((rows, cols))).astype(f32) for _ in range(n)])
    if n >= 2:
        mov[-1] = 1.5                      # constant image: the sigma clamp
    if n >= 3:
        mov[-2] = fixed                    # identical image: similarity 0
    a = xo.ncc(fixed, mov)
    b = ref_slice.ncc(fixed, mov)
    assert a.tobytes() == b.tobytes(), (a, b)
    mask = (rng.random((rows, cols)) < rng.uniform(0.3, 0.95)).astype(np.uint8)
    mask.flat[:2] = 1                      # at least two pixels (N - 1 in the denominator)
    am = xo.ncc(fixed, mov, mask=mask)
    bm = ref_slice.ncc(fixed, mov, mask=mask)
    assert am.tobytes() == bm.tobytes(), (am, bm)


def test_pose_distribution_and_pre_compute_are_the_reference_code(xo):
    """RayCaster::distribute_xforms_among_cam_models (xregRayCastInterface.cpp:97-114) and RayCasterCPU::pre_compute
    (xregRayCastBaseCPU.cpp:128-158): camera-major replication; REPLACE fills with the default value unless background
    projections are in use, ACCUM keeps the content, background projections are copied per camera association."""
    rng = np.random.default_rng(2)
    poses = rng.standard_normal((5, 12)).astype(f32)
    for n_cams in (1, 3):
        a, ia = xo.distribute_xforms(poses, n_cams)
        b, ib = ref_slice.distribute_xforms(poses, n_cams)
        assert a.tobytes() == b.tobytes() and np.array_equal(ia, ib)
    cam_idx = np.array([0, 0, 1, 1, 2, 0], np.uint32)
    bgs = [rng.standard_normal((4, 6)).astype(f32) for _ in range(3)]
    for store in (0, 1):
        for bg in (None, bgs):
            for default in (0.0, 2.5):
                a = rng.standard_normal((6, 4, 6)).astype(f32)
                b = a.copy()
                xo.pre_compute(a, cam_idx, bg_projs=bg, store_method=store, default_bg=default)
                ref_slice.pre_compute(b, cam_idx, 3, bg_projs=bg, store_method=store, default_bg=default)
                assert a.tobytes() == b.tobytes(), (store, bg is not None, default)


def test_view_combination_is_the_reference_code(xo):
    """ImgSimMetric2DCombineMean::compute (xregImgSimMetric2DCombine.cpp:79-98): the oracle's xo_combine_mean and the
    product's host-side combine_mean (xreg_b200.regi) equal it bit for bit."""
    from xreg_b200 import regi

    rng = np.random.default_rng(8)
    for views, poses in ((1, 7), (2, 5), (3, 100), (5, 1)):
        v = rng.uniform(0, 1, (views, poses)).astype(f32)
        ref = ref_slice.combine(v, mean=True)
        assert xo.combine_mean(v).tobytes() == ref.tobytes()
        assert regi.combine_mean(v).tobytes() == ref.tobytes()
        acc = np.zeros(poses, f32)
        for k in range(views):
            acc = (acc + v[k]).astype(f32)
        assert ref_slice.combine(v, mean=False).tobytes() == acc.tobytes()


# ---- gradient-NCC and patch gradient-NCC (the headline metric): the reference's class code over chosen filters -------------
def _oracle_filters(xo):
    def sobel(img, dx, dy):
        gx, gy = xo.sobel(img)
        return gx if dx == 1 else gy

    return (lambda img, k: xo.gauss_blur(img, k)), sobel


def _cv2_filters():
    import cv2

    return (lambda img, k: cv2.GaussianBlur(img, (k, k), 0)), (lambda img, dx, dy: cv2.Sobel(img, -1, dx, dy))


@pytest.mark.parametrize("seed", range(12))
def test_oracle_gradient_metrics_equal_the_reference_class_code(xo, seed):
    """ImgSimMetric2DGradImgCPU::{allocate_resources, compute_sobel_grads}, ImgSimMetric2DGradNCCCPU::{allocate_resources,
    compute, process_mask} and ImgSimMetric2DPatchGradNCCCPU::{allocate_resources, compute, process_mask}
    (xregImgSimMetric2DGradImgCPU.cpp:32-102, ...GradNCCCPU.cpp:29-81, ...PatchGradNCCCPU.cpp:34-253,313-409) compiled from
    the reference's own lines on top of its NCC / patch-NCC classes; cv::GaussianBlur / cv::Sobel are call-outs.
    (1) With the oracle's restatement of the two filters installed, the oracle's gradient-NCC and patch gradient-NCC equal
        the reference classes bit for bit: everything except OpenCV's arithmetic is pinned.
    (2) With the REAL OpenCV installed (cv2), the reference classes give what an xReg build gives up to OpenCV's version:
        the oracle agrees to <= 2e-6 (its Gaussian is within 1 ulp of OpenCV's, its Sobel identical)."""
    fixed, mov, mask, radius, stride, _ = _metric_case(100 + seed)
    width = [0, 3, 5, 7][seed % 4]
    ref_slice.set_filters(*_oracle_filters(xo))
    a = xo.grad_ncc(fixed, mov, mask=mask, gauss_width=width, n_threads=1)
    b = ref_slice.grad_ncc(fixed, mov, mask=mask, gauss_width=width)
    assert a.tobytes() == b.tobytes(), (a, b)
    for kw in (dict(), dict(compute_mean=True), dict(mask_stats=True) if mask is not None else dict(weight_sims=False)):
        opts = xo.patch_opts(radius=radius, stride=stride, **kw)
        w = xo.patch_weights(fixed.shape[0], fixed.shape[1], opts, mask=mask) if (mask is not None and opts.use_mask_for_weighting) else None
        pa = xo.patch_grad_ncc(fixed, mov, opts, mask=mask, weights=w, gauss_width=width, n_threads=1)
        pb = ref_slice.patch_grad_ncc(fixed, mov, opts, mask=mask, gauss_width=width)
        assert pa.tobytes() == pb.tobytes(), (kw, pa, pb)
    try:
        ref_slice.set_filters(*_cv2_filters())
    except ImportError:
        return
    c = ref_slice.grad_ncc(fixed, mov, mask=mask, gauss_width=width)
    assert np.max(np.abs(a - c)) <= 2e-6, (a, c)
    opts = xo.patch_opts(radius=radius, stride=stride)
    w = xo.patch_weights(fixed.shape[0], fixed.shape[1], opts, mask=mask) if mask is not None else None
    pa = xo.patch_grad_ncc(fixed, mov, opts, mask=mask, weights=w, gauss_width=width, n_threads=1)
    pc = ref_slice.patch_grad_ncc(fixed, mov, opts, mask=mask, gauss_width=width)
    finite = pa < 1e30
    assert np.array_equal(finite, pc < 1e30) and np.max(np.abs(pa[finite] - pc[finite]), initial=0.0) <= 2e-6, (pa, pc)


@pytest.mark.parametrize("seed", range(12))
def test_oracle_ssd_equals_the_reference_class_code(xo, seed):
    """ImgSimMetric2DSSDCPU::{allocate_resources, compute, process_mask} + ApplyMaskToEigenMatInPlace
    (xregImgSimMetric2DSSDCPU.cpp:29-109): masking, difference, division by the pixel count are the reference's lines; the one
    Eigen reduction follows the stated convention (as for unmasked NCC)."""
    rng = np.random.default_rng(3000 + seed)
    rows, cols = int(rng.integers(1, 40)), int(rng.integers(1, 60))
    n = int(rng.integers(1, 4))
    fixed = (rng.standard_normal((rows, cols)) * 3 + 5).astype(f32)
    mov = np.stack([(fixed + rng.standard_normal((rows, cols)) * rng.uniform(0, 2)).astype(f32) for _ in range(n)])
    mov[0] = fixed
    assert xo.ssd(fixed, mov).tobytes() == ref_slice.ssd(fixed, mov).tobytes()
    mask = (rng.random((rows, cols)) < 0.7).astype(np.uint8)
    a, b = xo.ssd(fixed, mov, mask), ref_slice.ssd(fixed, mov, mask)
    assert a.tobytes() == b.tobytes() and a[0] == 0


def test_product_exp_se3_matches_the_reference_code():
    """xrc_exp_se3 (host-only code of the product library; the pose composition of xrc_obj_fn_se3) against the reference's
    own SkewMatrix / WedgeSkew / ExpSO3 / ExpSE3 lines (xregRotUtils.cpp:33-105, xregRigidUtils.cpp:40-85): same SE(3)
    element to f32 rounding, from tiny to large rotations, and the Python host model too."""
    import ctypes as C

    from xreg_b200 import _lib
    from xreg_b200.geometry import exp_se3

    lib = _lib.load()
    FP = C.POINTER(C.c_float)
    rng = np.random.default_rng(5)
    cases = [np.zeros(6), [0, 0, 0, 3, -4, 5], [1e-9, 0, 0, 1, 2, 3], [0.3, -0.2, 0.1, 10, -20, 30], [3.0, 0.5, -0.4, 1, 1, 1]]
    cases += [np.concatenate([rng.normal(0, s, 3), rng.normal(0, 50, 3)]) for s in (1e-4, 0.05, 0.5, 1.5) for _ in range(10)]
    worst = 0.0
    for x in cases:
        x32 = np.asarray(x, f32)
        ref = ref_slice.exp_se3(x32).reshape(3, 4)
        out = np.zeros(12, f32)
        lib.xrc_exp_se3(x32.ctypes.data_as(FP), out.ctypes.data_as(FP))
        got = out.reshape(3, 4)
        scale = max(1.0, float(np.abs(ref[:, 3]).max()))
        worst = max(worst, float(np.abs(got[:, :3] - ref[:, :3]).max()), float(np.abs(got[:, 3] - ref[:, 3]).max()) / scale)
        model = exp_se3(x32.astype(np.float64))[:3]
        assert np.abs(model[:, :3] - ref[:, :3]).max() <= 2e-6 and np.abs(model[:, 3] - ref[:, 3]).max() <= 2e-6 * scale
        R = ref[:, :3].astype(np.float64)
        assert np.abs(R @ R.T - np.eye(3)).max() <= 5e-6
    assert worst <= 2e-6, worst


def _cam_bytes(c):
    return (int(c.rows), int(c.cols), bytes(c.intrins_inv), bytes(c.extrins_inv), bytes(c.pinhole), float(c.focal_len),
            int(c.frame_type))


def test_camera_setup_is_the_reference_code(xo):
    """CameraModel::setup (naive; intrinsics + extrinsics), MakeNaiveIntrins, FocalLenFromIntrins, SE3Inv and
    DownsampleCameraModel (xregPerspectiveXform.cpp:186-254,302-334,654-688; xregRigidUtils.cpp:29-38) compiled from the
    reference's lines: the oracle's camera set-up and the product's host mirror (xreg_b200.geometry: what the tests and the
    bench build their cameras with) produce the same cameras bit for bit, for all three frame types.  Eigen's 3x3 inverse
    and `-1 * R^T * t` are conventions shared by stand-in and oracle; everything around them is the reference's."""
    from xreg_b200.geometry import CameraModel as PyCam, downsample_camera_model

    rng = np.random.default_rng(12)
    for frame in (0, 1, 2):
        for (f, nr, nc, rs, cs) in ((1020.0, 1536, 1536, 0.194, 0.194), (400.0, 80, 96, 1.6, 1.5), (655.5, 7, 1001, 0.31, 2.7)):
            a = xo.cam_setup_naive(f, nr, nc, rs, cs, frame_type=frame)
            b = ref_slice.cam_setup_naive(f, nr, nc, rs, cs, frame_type=frame)
            assert _cam_bytes(a) == _cam_bytes(b)
            py = PyCam(coord_frame_type=frame).setup(f, nr, nc, rs, cs)
            assert _cam_bytes(xo.cam_struct(py)) == _cam_bytes(b)
        for _ in range(6):
            E = np.eye(4)
            E[:3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
            E[:3, 3] = rng.uniform(-300, 300, 3)
            K = np.array([[rng.uniform(-6000, -500), rng.uniform(-1, 1), rng.uniform(10, 800)],
                          [0, rng.uniform(-6000, -500), rng.uniform(10, 800)], [0, 0, 1]])
            nr, nc = int(rng.integers(8, 2000)), int(rng.integers(8, 2000))
            rs, cs = float(rng.uniform(0.1, 2.0)), float(rng.uniform(0.1, 2.0))
            a = xo.cam_setup(K, E, nr, nc, rs, cs, frame_type=frame)
            b = ref_slice.cam_setup(K, E, nr, nc, rs, cs, frame_type=frame)
            assert _cam_bytes(a) == _cam_bytes(b)
            src = PyCam(coord_frame_type=frame).setup_intrins_extrins(K.astype(f32), E.astype(f32), nr, nc, rs, cs)
            assert _cam_bytes(xo.cam_struct(src)) == _cam_bytes(b)
            for ds, even in ((0.5, False), (0.125, True), (0.3125, False), (0.25, True)):
                rb, kb, sp = ref_slice.cam_downsample(K, E, nr, nc, rs, cs, frame, ds, even)
                d = downsample_camera_model(src, ds, even)
                assert _cam_bytes(xo.cam_struct(d)) == _cam_bytes(rb), (frame, ds, even)
                assert np.asarray(d.intrins, f32).tobytes() == kb.tobytes()
                assert f32(d.det_row_spacing) == sp[0] and f32(d.det_col_spacing) == sp[1]


def test_volume_geometry_is_the_reference_code():
    """ITKImageIndexBoundsAsEigen / ITKImagePhysicalPointTransformsAsEigen (xregITKBasicImageUtils.h:54-75,131-168): the
    product's host mirror (Volume.idx_to_phys: what every test and the bench hand to xrc_rc_set_volumes) builds the same
    index -> physical transform bit for bit -- the double product Dir * spacing narrowed once to float -- and the index
    bounds are [0, size - 1]."""
    from xreg_b200.geometry import Volume, exp_se3

    rng = np.random.default_rng(21)
    for k in range(20):
        dims = [int(rng.integers(1, 700)) for _ in range(3)]
        spacing = rng.uniform(0.1, 3.0, 3)
        origin = rng.uniform(-500, 500, 3)
        D = exp_se3(np.concatenate([rng.normal(0, 0.7, 3), np.zeros(3)]))[:3, :3].astype(np.float64) if k % 2 else np.eye(3)
        mn, mx, a = ref_slice.itk_volume_geometry(dims, origin, spacing, D)
        vol = Volume(np.zeros((1, 1, 1), f32), spacing=tuple(spacing), origin=tuple(origin), direction=D)
        assert np.asarray(vol.idx_to_phys(), f32).tobytes() == a.tobytes()
        assert np.array_equal(mn, np.zeros(3, f32)) and np.array_equal(mx, (np.array(dims) - 1).astype(f32))


# ---- log remap of a projection -------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(10))
def test_oracle_log_remap_equals_the_reference_filter_code(xo, seed):
    """ImageIntensLogTransFilter::GenerateData (lib/image/xregImageIntensLogTrans.cpp:55-144), the reference's own lines,
    against xo_log_remap, bit for bit, in its three modes: normalise to [0, 1]; I0 = max of the smoothed image (the
    DiscreteGaussianImageFilter is a call-out: ITK is absent, the oracle's restatement of it is installed -- that piece
    is unpinned); a given I0.  Zeros and sub-eps pixels take the value of the smallest positive pixel."""
    rng = np.random.default_rng(500 + seed)
    rows, cols = int(rng.integers(3, 60)), int(rng.integers(3, 70))
    img = (rng.uniform(0.0, 4000.0, (rows, cols)) * (rng.random((rows, cols)) < 0.9)).astype(f32)
    img[rng.integers(0, rows), rng.integers(0, cols)] = f32(5.0e-7)     # below eps
    if seed % 3 == 0:
        img *= f32(1.0e-3)
    calls = []

    def gauss(a, var):
        calls.append(var)
        return xo.itk_discrete_gaussian_2d(a, var)

    for norm, use_max, i0 in ((False, True, 1.0), (True, True, 1.0), (False, False, 4096.0), (True, False, 2.5)):
        want = ref_slice.log_remap(img, norm, use_max, i0, gaussian=gauss)
        got, _ = xo.log_remap(img, norm, use_max, i0)
        assert got.tobytes() == want.tobytes(), (norm, use_max, i0)
        assert np.all(np.isfinite(got)) and got.max() == got[img <= 1.0e-6].max()
    assert calls == [2.0]      # only the default mode smooths, with variance 2 (:100)
