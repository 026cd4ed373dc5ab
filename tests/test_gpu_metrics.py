"""GPU parity of ImgSimMetric2D{NCC,GradNCC,PatchNCC,PatchGradNCC}CUDA against the CPU oracle:
similarity values within 1e-5 absolute, gradient images bit-exact."""
import numpy as np
import pytest

import xreg_b200
from xreg_b200 import synth
from xreg_b200.geometry import to12

pytestmark = pytest.mark.gpu
f32 = np.float32
SIM_TOL = 1.0e-5  # BASELINE.json north_star: similarity values within 1e-5 absolute


def _img(rows, cols, seed, smooth=True):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((rows, cols))
    if smooth:
        k = np.ones(5) / 5
        if rows >= 5:
            a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 0, a)
        if cols >= 5:
            a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, a)
    return (a * 3 + 5).astype(f32)


def _movs(fixed, n, seed):
    out = [(_img(*fixed.shape, seed + i) * (0.2 + 0.1 * i) + (1.0 - 0.1 * i) * fixed).astype(f32) for i in range(n)]
    out[-1] = np.full_like(fixed, 2.5)  # constant image: sigma clamp path
    return np.ascontiguousarray(np.stack(out))


def _run(sm, fixed, mov, mask=None):
    sm.set_num_moving_images(mov.shape[0])
    sm.set_fixed_image(fixed)
    sm.set_mov_imgs_host_buf(mov)
    if mask is not None:
        sm.set_mask(mask)
    sm.allocate_resources()
    sm.compute()
    return sm.sim_vals().copy()


SHAPES = [(61, 73), (32, 32), (100, 37), (7, 9)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("masked", [False, True])
def test_ncc(ctx, xo, shape, masked):
    fixed = _img(*shape, seed=1)
    mov = _movs(fixed, 5, seed=10)
    mask = (np.random.default_rng(2).random(shape) > 0.3).astype(np.uint8) if masked else None
    got = _run(xreg_b200.ImgSimMetric2DNCCCUDA(ctx), fixed, mov, mask)
    ref = xo.ncc(fixed, mov, mask=mask)
    assert np.max(np.abs(got - ref)) <= SIM_TOL
    assert abs(got[-1] - 0.5) < 1e-6


@pytest.mark.parametrize("shape", SHAPES + [(480, 480)])
@pytest.mark.parametrize("masked", [False, True])
def test_ssd(ctx, xo, shape, masked):
    """ImgSimMetric2DSSDCUDA vs ImgSimMetric2DSSDCPU's restatement: SSD is not normalised, so the tolerance is
    relative (1e-5, the rounding noise of the reference's own f32 sum)."""
    fixed = _img(*shape, seed=21)
    mov = _movs(fixed, 4, seed=22)
    mov[1] = fixed                                   # identical image: exactly 0
    mask = (np.random.default_rng(23).random(shape) > 0.35).astype(np.uint8) if masked else None
    got = _run(xreg_b200.ImgSimMetric2DSSDCUDA(ctx), fixed, mov, mask)
    ref = xo.ssd(fixed, mov, mask)
    assert ref[1] == 0.0 and abs(got[1]) <= 1e-9 * ref.max()   # f64 moments: the cancellation leaves ~1e-12 relative
    assert np.all(np.abs(got - ref) <= 1e-5 * ref + 1e-9 * ref.max())


@pytest.mark.parametrize("shape", SHAPES + [(64, 96), (3, 40)])
@pytest.mark.parametrize("width", [0, 3, 5, 7, 9])
def test_gradient_images_bit_exact(ctx, xo, shape, width):
    fixed = _img(*shape, seed=3, smooth=False)
    mov = _movs(fixed, 3, seed=20)
    sm = xreg_b200.ImgSimMetric2DGradNCCCUDA(ctx)
    sm.set_smooth_img_before_sobel_kernel_radius(width)
    _run(sm, fixed, mov)
    for i in range(3):
        gx, gy = sm.read_grads(i)
        rx, ry = xo.grad_imgs(mov[i], width)
        np.testing.assert_array_equal(gx, rx)
        np.testing.assert_array_equal(gy, ry)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("masked", [False, True])
def test_grad_ncc(ctx, xo, shape, masked):
    fixed = _img(*shape, seed=4)
    mov = _movs(fixed, 4, seed=30)
    mask = (np.random.default_rng(5).random(shape) > 0.25).astype(np.uint8) if masked else None
    sm = xreg_b200.ImgSimMetric2DGradNCCCUDA(ctx)
    got = _run(sm, fixed, mov, mask)
    ref = xo.grad_ncc(fixed, mov, mask=mask, gauss_width=5)
    assert np.max(np.abs(got - ref)) <= SIM_TOL
    sm0 = xreg_b200.ImgSimMetric2DGradNCCCUDA(ctx)
    sm0.set_smooth_img_before_sobel_kernel_radius(0)
    got0 = _run(sm0, fixed, mov, mask)
    assert np.max(np.abs(got0 - xo.grad_ncc(fixed, mov, mask=mask, gauss_width=0))) <= SIM_TOL


@pytest.mark.parametrize("shape,radius,stride", [((61, 73), 5, 1), ((40, 300), 10, 1), ((33, 35), 3, 2),
                                                 ((11, 11), 5, 1), ((64, 64), 13, 3), ((50, 520), 21, 1)])
def test_patch_ncc(ctx, xo, shape, radius, stride):
    fixed = _img(*shape, seed=6)
    mov = _movs(fixed, 3, seed=40)
    sm = xreg_b200.ImgSimMetric2DPatchNCCCUDA(ctx)
    sm.set_patch_radius(radius)
    sm.set_patch_stride(stride)
    got = _run(sm, fixed, mov)
    ref = xo.patch_ncc(fixed, mov, xo.patch_opts(radius=radius, stride=stride))
    assert np.max(np.abs(got - ref)) <= SIM_TOL
    assert abs(got[-1] - 1.0) < 1e-6
    assert sm.num_patches() == xo.num_patches(shape[0], shape[1], radius, stride)


@pytest.mark.parametrize("mode", ["mean", "unweighted", "mask", "mask_stats", "wgt_img"])
def test_patch_ncc_options(ctx, xo, mode):
    shape = (45, 52)
    fixed = _img(*shape, seed=7)
    mov = _movs(fixed, 3, seed=50)
    mask = np.zeros(shape, np.uint8)
    mask[5:40, 8:45] = 1
    mask[20:24, 20:30] = 0
    sm = xreg_b200.ImgSimMetric2DPatchNCCCUDA(ctx)
    sm.set_patch_radius(4)
    o = xo.patch_opts(radius=4)
    use_mask, wgt_img = None, None
    if mode == "mean":
        sm.set_compute_mean_of_patch_sims(True)
        o.compute_mean_of_patch_sims = 1
    elif mode == "unweighted":
        sm.set_weight_patch_sims_in_combine(False)
        o.weight_patch_sims = 0
    elif mode == "mask":
        use_mask = mask
    elif mode == "mask_stats":
        use_mask = mask
        sm.set_use_mask_for_patch_stats(True)
        o.use_mask_for_patch_stats = 1
    elif mode == "wgt_img":
        wgt_img = np.random.default_rng(8).random(shape).astype(f32)
        sm.set_wgt_img(wgt_img)
    got = _run(sm, fixed, mov, use_mask)
    w = xo.patch_weights(shape[0], shape[1], o, mask=use_mask, wgt_img=wgt_img) if (use_mask is not None or wgt_img is not None) else None
    ref = xo.patch_ncc(fixed, mov, o, mask=use_mask, weights=w)
    tol = SIM_TOL if mode != "unweighted" else SIM_TOL * ref.max()
    assert np.max(np.abs(got - ref)) <= tol


@pytest.mark.parametrize("shape,radius", [((61, 73), 5), ((96, 96), 10), ((48, 300), 13)])
@pytest.mark.parametrize("masked", [False, True])
def test_patch_grad_ncc(ctx, xo, shape, radius, masked):
    fixed = _img(*shape, seed=9)
    mov = _movs(fixed, 3, seed=60)
    mask = synth.circular_mask(*shape) if masked else None
    sm = xreg_b200.ImgSimMetric2DPatchGradNCCCUDA(ctx)
    sm.set_patch_radius(radius)
    got = _run(sm, fixed, mov, mask)
    o = xo.patch_opts(radius=radius)
    w = xo.patch_weights(shape[0], shape[1], o, mask=mask) if masked else None
    ref = xo.patch_grad_ncc(fixed, mov, o, mask=mask, weights=w, gauss_width=5)
    assert np.max(np.abs(got - ref)) <= SIM_TOL


def test_patch_grad_ncc_on_drrs_with_flat_regions(ctx, xo, small_scene):
    """DRRs have exactly-zero background: the sigma clamp / zero-variance patches must agree."""
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, 4)
    rc = xreg_b200.RayCasterLineIntCUDA(ctx)
    rc.set_volume(vol)
    rc.set_camera_model(cam)
    rc.set_num_projs(4)
    rc.allocate_resources()
    rc.set_xforms_cam_to_itk_phys(list(poses))
    rc.compute()
    drrs = rc.raw_host_pixel_buf()
    fixed = synth.add_noise(drrs[0])
    for cls, fn in ((xreg_b200.ImgSimMetric2DPatchGradNCCCUDA, lambda: xo.patch_grad_ncc(fixed, drrs, xo.patch_opts(radius=5))),
                    (xreg_b200.ImgSimMetric2DGradNCCCUDA, lambda: xo.grad_ncc(fixed, drrs)),
                    (xreg_b200.ImgSimMetric2DNCCCUDA, lambda: xo.ncc(fixed, drrs)),
                    (xreg_b200.ImgSimMetric2DPatchNCCCUDA, lambda: xo.patch_ncc(fixed, drrs, xo.patch_opts(radius=5)))):
        sm = cls(ctx)
        sm.set_num_moving_images(4)
        sm.set_fixed_image(fixed)
        sm.set_mov_imgs_buf_from_ray_caster(rc)
        sm.allocate_resources()
        sm.compute()
        assert np.max(np.abs(sm.sim_vals() - fn())) <= SIM_TOL, cls.__name__
        assert int(np.argmin(sm.sim_vals())) == 0


def test_rebinding_offsets_num_images_and_mask_update(ctx, xo):
    fixed = _img(40, 44, seed=11)
    mov = _movs(fixed, 6, seed=70)
    sm = xreg_b200.ImgSimMetric2DGradNCCCUDA(ctx)
    sm.set_num_moving_images(6)
    sm.set_fixed_image(fixed)
    sm.set_mov_imgs_host_buf(mov)
    sm.allocate_resources()
    sm.compute()
    full = sm.sim_vals().copy()
    # fewer images at an offset (xregImgSimMetric2DCPU.cpp:59-65)
    sm.set_num_moving_images(2)
    sm.set_mov_imgs_host_buf(mov, 3)
    sm.compute()
    np.testing.assert_array_equal(sm.sim_vals(), full[3:5])
    with pytest.raises(xreg_b200.XregError):
        sm.set_num_moving_images(7)
    # mask set after allocation is picked up lazily (process_updated_mask)
    mask = synth.circular_mask(40, 44)
    sm.set_mask(mask)
    sm.set_num_moving_images(6)
    sm.set_mov_imgs_host_buf(mov, 0)
    sm.compute()
    assert np.max(np.abs(sm.sim_vals() - xo.grad_ncc(fixed, mov, mask=mask))) <= SIM_TOL
    sm.set_mask(None)
    sm.compute()
    np.testing.assert_array_equal(sm.sim_vals(), full)


def test_metric_errors(ctx):
    sm = xreg_b200.ImgSimMetric2DPatchGradNCCCUDA(ctx)
    with pytest.raises(xreg_b200.XregError):
        sm.compute()
    sm.set_num_moving_images(1)
    sm.set_fixed_image(_img(8, 8, 0))
    with pytest.raises(xreg_b200.XregError):
        sm.allocate_resources()  # nothing bound
    sm.set_mov_imgs_host_buf(_movs(_img(8, 8, 0), 1, 1))
    sm.set_patch_radius(5)
    with pytest.raises(xreg_b200.XregError):
        sm.allocate_resources()  # patch diameter 11 > 8 (xregImgSimMetric2DPatchCommon.cpp:272-273)
    with pytest.raises(xreg_b200.XregError):
        sm.set_smooth_img_before_sobel_kernel_radius(4)  # width must be odd
    sm.set_choose_rand_patches(True)          # supported since round 2 (test_patch_subsets_follow_the_reference)
    sm.set_choose_rand_patches(False)


def test_combine_mean(ctx, xo):
    fixed = _img(30, 30, seed=12)
    mov = _movs(fixed, 4, seed=80)
    a = xreg_b200.ImgSimMetric2DNCCCUDA(ctx)
    b = xreg_b200.ImgSimMetric2DGradNCCCUDA(ctx)
    _run(a, fixed, mov)
    _run(b, fixed, mov)
    comb = xreg_b200.ImgSimMetric2DCombineMean()
    comb.set_sim_metrics([a, b])
    comb.compute()
    np.testing.assert_allclose(comb.sim_vals(), xo.combine_mean(np.stack([a.sim_vals(), b.sim_vals()])), rtol=1e-6)


# ---- reference-order combine of the per-patch values (XRC_COMBINE_*) -------------------------------------------

def _seqsum(ctx, seqs, serial=False):
    import ctypes as C

    from xreg_b200 import _lib

    lib = _lib.load()
    v = np.ascontiguousarray(seqs, dtype=f32)
    out = np.zeros(v.shape[0], dtype=f32)
    FP = C.POINTER(C.c_float)
    _lib.check(lib.xrc_seqsum_f32(ctx.handle, v.ctypes.data_as(FP), v.shape[0], v.shape[1], int(serial),
                                  out.ctypes.data_as(FP)))
    return out


def _seq_literal(seqs):
    """the reference's loop: `Scalar sum = 0; for (s : vals) sum += s;` (np.cumsum on float32 is sequential)"""
    v = np.ascontiguousarray(seqs, dtype=f32)
    if v.shape[1] == 0:
        return np.zeros(v.shape[0], dtype=f32)
    with np.errstate(all="ignore"):
        return np.cumsum(v, axis=1, dtype=f32)[:, -1]


def _bits(a):
    return np.ascontiguousarray(a, dtype=f32).view(np.uint32)


@pytest.mark.parametrize("n", [0, 1, 5, 127, 128, 129, 1023, 4099, 16384, 16385, 206116, 557000, 1600000])
def test_seqsum_emulation_is_bit_exact(ctx, n):
    """patch_seqsum_kernel == the literal sequential f32 loop, bit for bit, on the sequences the metric produces
    (per-patch values in [0, 2], near-constant weighted values) and on adversarial ones (ties at every step, a sum
    walking down across a binade, mixed signs, tiny and huge magnitudes, NaN)."""
    rng = np.random.default_rng(n + 1)
    seqs = [
        rng.random(n),                                         # per-patch similarities, weights 1
        rng.random(n) * 3.0e-4,                                # normalised weights times similarities
        np.full(n, 1.0 / 3800.0),                              # full-coverage weights (total-weight sum)
        rng.standard_normal(n),                                # mixed signs: the sum keeps crossing zero
        -rng.random(n),                                        # negative running sum
        rng.random(n) * 1.0e-20,
        np.concatenate([[1.0e30], rng.random(max(n - 1, 0)) * 1.0e29])[:n],   # leaves the fast range
        np.concatenate([[1.0], np.full(max(n - 1, 0), 2.0 ** -24)])[:n],      # every addend is a tie
        np.concatenate([[8388608.0], np.full(max(n - 1, 0), -0.75)])[:n],     # walks down out of its binade
        (rng.random(n) * 2 - 0.3) * 1.0e-38,                   # denormal range
        np.where(rng.random(n) < 0.5, 0.0, rng.random(n)),     # skipped patches contribute +0
    ]
    v = np.stack([np.asarray(s, dtype=f32) for s in seqs])
    want = _seq_literal(v)
    np.testing.assert_array_equal(_bits(_seqsum(ctx, v, serial=True)), _bits(want))
    np.testing.assert_array_equal(_bits(_seqsum(ctx, v)), _bits(want))             # two-phase emulation (default)
    np.testing.assert_array_equal(_bits(_seqsum(ctx, v, serial=2)), _bits(want))   # chained emulation (fallback)
    for cluster in (1, 2, 4, 8):                                                   # two-phase, forced cluster size
        np.testing.assert_array_equal(_bits(_seqsum(ctx, v, serial=10 + cluster)), _bits(want))
    if n > 200:
        w = v[:1].copy()
        w[0, n // 2] = np.nan
        assert np.isnan(_seqsum(ctx, w)[0])


def _drr_scene_images(ctx, xo, small_scene, n=4):
    vol, cam, nominal = small_scene
    poses = synth.pose_population(vol, nominal, n)
    drrs = xo.drr(vol.data, vol.idx_to_phys(), [xo.cam_struct(cam)], to12(poses))
    return synth.add_noise(drrs[0]), np.ascontiguousarray(drrs)


@pytest.mark.parametrize("kind", ["patch-ncc", "patch-grad-ncc"])
@pytest.mark.parametrize("frac", [None, 0.45, 0.9])
def test_patch_combine_follows_the_reference_sum(ctx, xo, small_scene, kind, frac):
    """With mask-coverage weights the reference's sequential f32 sums carry a systematic error of a few 1e-5
    (oracle vs an exact float64 model, checked below); the default combine mode reproduces that sum, so the CUDA
    value agrees with the CPU class to ~1e-6; the literal-loop mode is bitwise identical to it; the f64 mode
    agrees with the exact model instead."""
    from tests.helpers import patch_ncc_model_f64

    fixed, drrs = _drr_scene_images(ctx, xo, small_scene)
    rows, cols = fixed.shape
    mask = synth.circular_mask(rows, cols, frac) if frac else None
    grad = kind == "patch-grad-ncc"
    cls = xreg_b200.ImgSimMetric2DPatchGradNCCCUDA if grad else xreg_b200.ImgSimMetric2DPatchNCCCUDA
    o = xo.patch_opts(radius=5)
    w = xo.patch_weights(rows, cols, o, mask=mask) if mask is not None else None
    ref = (xo.patch_grad_ncc(fixed, drrs, o, mask=mask, weights=w) if grad else xo.patch_ncc(fixed, drrs, o, mask=mask, weights=w))

    got = {}
    for mode in ("reference", "reference-serial", "f64"):
        sm = cls(ctx)
        sm.set_patch_radius(5)
        sm.set_combine_mode(mode)
        got[mode] = _run(sm, fixed, drrs, mask)
    np.testing.assert_array_equal(_bits(got["reference"]), _bits(got["reference-serial"]))
    assert np.max(np.abs(got["reference"] - ref)) <= 2.0e-6

    # exact float64 model of the same formula on the oracle's gradient images
    def exact(k):
        if not grad:
            return patch_ncc_model_f64(fixed, drrs[k], 5, mask=mask, weights=w)
        fgx, fgy = xo.grad_imgs(fixed, 5)
        gx, gy = xo.grad_imgs(drrs[k], 5)
        return 0.5 * (patch_ncc_model_f64(fgx, gx, 5, mask=mask, weights=w) + patch_ncc_model_f64(fgy, gy, 5, mask=mask, weights=w))

    ex = np.array([exact(k) for k in range(2)])
    assert np.max(np.abs(got["f64"][:2] - ex)) <= 2.0e-6
    if frac == 0.9:
        # the point of the default mode: here the reference itself is > 1e-5 away from exact arithmetic
        assert np.max(np.abs(ref[:2] - ex)) > 1.0e-5


def test_combine_mode_switch_after_allocation(ctx, xo):
    fixed = _img(61, 73, seed=3)
    mov = _movs(fixed, 3, seed=90)
    sm = xreg_b200.ImgSimMetric2DPatchNCCCUDA(ctx)
    sm.set_patch_radius(4)
    sm.set_combine_mode("f64")
    a = _run(sm, fixed, mov)
    sm.set_combine_mode("reference")   # per-patch buffer is allocated on demand
    sm.compute()
    b = sm.sim_vals().copy()
    ref = xo.patch_ncc(fixed, mov, xo.patch_opts(radius=4))
    assert np.max(np.abs(b - ref)) <= 2.0e-6 and np.max(np.abs(a - ref)) <= SIM_TOL
    with pytest.raises(KeyError):
        sm.set_combine_mode("bogus")


@pytest.mark.parametrize("grad", [False, True])
@pytest.mark.parametrize("mode", ["default", "mean", "unweighted", "mask"])
def test_patch_subsets_follow_the_reference(ctx, xo, grad, mode):
    """set_patches_to_use / reset_patches_to_use and random patches (xregImgSimMetric2DPatchCommon.cpp:231-241,
    413-493): the metric over a local patch list -- unordered, with repeats, a single patch -- equals the oracle's
    xo_patch_ncc_subset (itself bit-equal to the reference's class code with set_patches_to_use), the list can change
    between computes without re-allocation, and resetting it restores the whole grid bit for bit."""
    shape = (57, 66)
    fixed = _img(*shape, seed=21)
    mov = _movs(fixed, 4, seed=70)
    mask = synth.circular_mask(*shape, 0.8) if mode == "mask" else None
    cls = xreg_b200.ImgSimMetric2DPatchGradNCCCUDA if grad else xreg_b200.ImgSimMetric2DPatchNCCCUDA
    sm = cls(ctx)
    sm.set_patch_radius(5)
    sm.set_patch_stride(2)
    o = xo.patch_opts(radius=5, stride=2)
    if mode == "mean":
        sm.set_compute_mean_of_patch_sims(True)
        o.compute_mean_of_patch_sims = 1
    elif mode == "unweighted":
        sm.set_weight_patch_sims_in_combine(False)
        o.weight_patch_sims = 0
    full = _run(sm, fixed, mov, mask)
    n_p = sm.num_patches()
    assert n_p == xo.num_patches(shape[0], shape[1], 5, 2)
    w = xo.patch_weights(shape[0], shape[1], o, mask=mask) if mask is not None else None
    gw = 5 if grad else None
    rng = np.random.default_rng(5)
    for sub in (rng.integers(0, n_p, size=n_p // 4), np.array([n_p - 1]), np.repeat(rng.integers(0, n_p, size=5), 3),
                np.arange(n_p)[::-1]):
        sm.set_patches_to_use(sub)
        sm.compute()
        got = sm.sim_vals().copy()
        ref = xo.patch_ncc_subset(fixed, mov, o, sub, mask=mask, weights=w, gauss_width=gw)
        # a list whose total weight is zero (one patch outside the mask) divides 0 by 0 in the reference too
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (mode, len(sub))
        ok = ~np.isnan(ref)
        tol = SIM_TOL if mode != "unweighted" else SIM_TOL * max(1.0, float(np.max(np.abs(ref[ok]), initial=0.0)))
        assert np.max(np.abs(got[ok] - ref[ok]), initial=0.0) <= tol, (mode, len(sub))
    sm.reset_patches_to_use()
    sm.compute()
    np.testing.assert_array_equal(sm.sim_vals(), full)
    # random patches: a fresh weighted draw per compute(), separated by the minimum distance
    sm.seed_rand_patches(11)
    sm.set_choose_rand_patches(True)
    sm.set_num_rand_patches(30)
    draws = []
    for _ in range(2):
        sm.compute()
        inds = sm.patch_inds_to_use().copy()
        assert inds.size == 30
        ref = xo.patch_ncc_subset(fixed, mov, o, inds, mask=mask, weights=w, gauss_width=gw)
        tol = SIM_TOL if mode != "unweighted" else SIM_TOL * max(1.0, float(np.max(np.abs(ref))))
        assert np.max(np.abs(sm.sim_vals() - ref)) <= tol
        ncc_ = (shape[1] - 1 - 10) // 2 + 1
        c = np.stack([5 + (inds // ncc_) * 2, 5 + (inds % ncc_) * 2], axis=1).astype(np.float64)
        d = np.linalg.norm(c[:, None] - c[None], axis=2) + np.eye(30) * 1e9
        assert d.min() >= np.sqrt(2.0 * 25) - 1e-9
        if mask is not None:
            assert np.all(w[inds.astype(np.int64)] > 0)      # never a patch of weight zero
        draws.append(inds)
    assert not np.array_equal(draws[0], draws[1])
    sm.set_choose_rand_patches(False)
    sm.compute()
    np.testing.assert_array_equal(sm.sim_vals(), full)
    with pytest.raises(xreg_b200.XregError):
        sm.set_patches_to_use([n_p])
