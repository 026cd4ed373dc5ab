"""Pins for the CPU oracle's similarity metrics: known answers,
OpenCV (cv2) for the Gaussian / Sobel arithmetic, numpy float64 models.  CPU only."""
import numpy as np
import pytest

from tests.helpers import ncc_model_f64, patch_ncc_model_f64

f32 = np.float32


def _img(rows=61, cols=73, seed=0, smooth=True):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((rows, cols))
    if smooth:
        k = np.ones(5) / 5
        a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 0, a)
        a = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, a)
    return (a * 3 + 5).astype(f32)


def test_ncc_known_answers(xo):
    x = _img()
    n = x.size
    # A.4 item 6: perfectly correlated -> ncc = (N-1)/N
    s = xo.ncc(x, np.stack([2.0 * x + 1.0, -0.5 * x + 3.0, np.full_like(x, 4.0)]))
    assert abs(s[0] - 0.5 * (1 - (n - 1) / n)) < 2e-6
    assert abs(s[1] - 0.5 * (1 + (n - 1) / n)) < 2e-6
    assert abs(s[2] - 0.5) < 1e-6  # constant moving image: sigma clamps at 1e-6, ncc -> 0


def test_ncc_matches_float64_model_and_overwrites_moving(xo):
    f = _img(seed=1)
    m = np.stack([_img(seed=2), 0.3 * f + _img(seed=3)])
    mask = (np.random.default_rng(4).random(f.shape) > 0.3).astype(np.uint8)
    s = xo.ncc(f, m)
    sm = xo.ncc(f, m, mask=mask)
    for k in range(2):
        assert abs(s[k] - ncc_model_f64(f, m[k])) < 2e-6
        assert abs(sm[k] - ncc_model_f64(f, m[k], mask)) < 2e-6
    buf = m.copy()
    xo.ncc(f, buf, inplace=True)
    assert abs(buf[0].mean()) < 1e-5  # zero-mean in place (xregImgSimMetric2DNCCCPU.h:36)


def test_ssd_known_answers_and_float64_model(xo):
    """ImgSimMetric2DSSDCPU (xregImgSimMetric2DSSDCPU.cpp:62-110): sum((f - m)^2) / num_pixels over the whole image,
    both images zero outside the mask."""
    f = _img(seed=5)
    n = f.size
    assert xo.ssd(f, f[None])[0] == 0.0
    assert abs(xo.ssd(f, (f + f32(2.0))[None])[0] - 4.0) < 1e-5            # constant offset c -> c^2
    m = np.stack([_img(seed=6), 0.5 * f])
    s = xo.ssd(f, m)
    for k in range(2):
        ref = np.sum((f.astype(np.float64) - m[k].astype(np.float64)) ** 2) / n
        assert abs(s[k] - ref) <= 1e-5 * ref
    mask = (np.random.default_rng(7).random(f.shape) > 0.4).astype(np.uint8)
    sm = xo.ssd(f, m, mask=mask)
    for k in range(2):
        d = (f.astype(np.float64) - m[k].astype(np.float64)) * mask
        assert abs(sm[k] - np.sum(d ** 2) / n) <= 1e-5 * sm[k]              # the divisor stays the full pixel count


def test_gauss_kernel_tables(xo):
    cv2 = pytest.importorskip("cv2")
    for k in (1, 3, 5, 7):
        np.testing.assert_array_equal(xo.gauss_kernel(k), cv2.getGaussianKernel(k, 0, cv2.CV_32F).ravel())
    np.testing.assert_array_equal(xo.gauss_kernel(5), np.array([1, 4, 6, 4, 1], dtype=f32) / 16)
    assert abs(xo.gauss_kernel(9).sum() - 1) < 1e-6


def test_sobel_bit_exact_vs_cv2(xo):
    cv2 = pytest.importorskip("cv2")
    for seed, shape in ((0, (61, 73)), (1, (8, 5)), (2, (2, 2)), (3, (1, 9))):
        img = _img(*shape, seed=seed, smooth=False)
        gx, gy = xo.sobel(img)
        if shape[1] >= 16:
            np.testing.assert_array_equal(gx, cv2.Sobel(img, -1, 1, 0))
            np.testing.assert_array_equal(gy, cv2.Sobel(img, -1, 0, 1))
        else:
            # cv2's scalar code path for very narrow images adds in a different order (1 ulp)
            np.testing.assert_allclose(gx, cv2.Sobel(img, -1, 1, 0), rtol=3e-7, atol=2e-6)
            np.testing.assert_allclose(gy, cv2.Sobel(img, -1, 0, 1), rtol=3e-7, atol=2e-6)


def test_sobel_of_ramp(xo):
    # A.4 item 7: interior = 8 * slope, reflected border columns / rows = 0
    r, c = np.meshgrid(np.arange(20, dtype=f32), np.arange(30, dtype=f32), indexing="ij")
    gx, gy = xo.sobel((0.5 * c + 0.25 * r).astype(f32))
    assert np.all(gx[:, 1:-1] == 4.0) and np.all(gx[:, [0, -1]] == 0)
    assert np.all(gy[1:-1, :] == 2.0) and np.all(gy[[0, -1], :] == 0)


def test_gaussian_blur_vs_cv2(xo):
    cv2 = pytest.importorskip("cv2")
    for k in (3, 5, 7):
        img = _img(seed=k, smooth=False)
        ref = cv2.GaussianBlur(img, (k, k), 0, 0)
        out = xo.gauss_blur(img, k)
        # cv2 4.13's SIMD path may contract to FMA: 1 ulp at most
        assert np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)) < 2.5e-7
    np.testing.assert_array_equal(xo.gauss_blur(img, 0), img)
    # tiny image: multiple reflections
    tiny = _img(3, 2, seed=9, smooth=False)
    np.testing.assert_allclose(xo.gauss_blur(tiny, 5), cv2.GaussianBlur(tiny, (5, 5), 0, 0), rtol=3e-7)


def test_grad_imgs_vs_cv2(xo):
    cv2 = pytest.importorskip("cv2")
    img = _img(seed=11)
    gx, gy = xo.grad_imgs(img, 5)
    b = cv2.GaussianBlur(img, (5, 5), 0, 0)
    np.testing.assert_allclose(gx, cv2.Sobel(b, -1, 1, 0), atol=2e-5)
    np.testing.assert_allclose(gy, cv2.Sobel(b, -1, 0, 1), atol=2e-5)


def test_grad_ncc_composition(xo):
    f = _img(seed=20)
    m = np.stack([_img(seed=21), 0.7 * f + 0.2 * _img(seed=22)])
    s = xo.grad_ncc(f, m, gauss_width=5)
    fgx, fgy = xo.grad_imgs(f, 5)
    for k in range(2):
        gx, gy = xo.grad_imgs(m[k], 5)
        expect = 0.5 * (ncc_model_f64(fgx, gx) + ncc_model_f64(fgy, gy))
        assert abs(s[k] - expect) < 2e-6
    s0 = xo.grad_ncc(f, m, gauss_width=0)
    assert np.all(np.abs(s0 - s) > 1e-4)  # smoothing matters
    assert xo.grad_ncc(f, f[None])[0] < 1e-3


def test_patch_grid_and_weights(xo):
    assert xo.num_patches(20, 30, 5, 1) == (20 - 10) * (30 - 10)
    assert xo.num_patches(20, 30, 5, 3) == 4 * 7
    assert xo.num_patches(11, 11, 5, 1) == 1
    o = xo.patch_opts(radius=2, stride=2)
    mask = np.zeros((12, 14), np.uint8)
    mask[:, 7:] = 1
    w = xo.patch_weights(12, 14, o, mask=mask)
    assert w.size == xo.num_patches(12, 14, 2, 2)
    assert abs(w.sum() - 1) < 1e-6 and w.reshape(4, 5)[0, 0] == 0 and w.reshape(4, 5)[0, -1] > 0
    assert np.all(xo.patch_weights(12, 14, o) == 1)


@pytest.mark.parametrize("radius,stride", [(2, 1), (3, 2), (5, 1)])
def test_patch_ncc_matches_float64_model(xo, radius, stride):
    f = _img(31, 37, seed=30)
    m = np.stack([_img(31, 37, seed=31), 0.5 * f + 0.5 * _img(31, 37, seed=32)])
    o = xo.patch_opts(radius=radius, stride=stride)
    s, ps = xo.patch_ncc(f, m, o, want_patch_sims=True)
    for k in range(2):
        assert abs(s[k] - patch_ncc_model_f64(f, m[k], radius, stride)) < 5e-6
    # A.4 item 8: identical images -> 1 - (n-1)/n for every textured patch
    n = (2 * radius + 1) ** 2
    s_id, ps_id = xo.patch_ncc(f, f[None], o, want_patch_sims=True)
    assert np.max(np.abs(ps_id - (1 - (n - 1) / n))) < 5e-6
    # mean-of-patches and unweighted-sum variants
    om = xo.patch_opts(radius=radius, stride=stride, compute_mean=True)
    assert abs(xo.patch_ncc(f, m, om)[0] - patch_ncc_model_f64(f, m[0], radius, stride, mean=True)) < 5e-6
    ou = xo.patch_opts(radius=radius, stride=stride, weight_sims=False)
    assert abs(xo.patch_ncc(f, m, ou)[0] / ps.shape[1] - s[0]) < 1e-5


def test_patch_ncc_with_mask_and_weights(xo):
    f = _img(33, 35, seed=40)
    m = _img(33, 35, seed=41)[None]
    mask = np.zeros(f.shape, np.uint8)
    mask[4:30, 6:28] = 1
    o = xo.patch_opts(radius=3)
    w = xo.patch_weights(33, 35, o, mask=mask)
    s = xo.patch_ncc(f, m, o, mask=mask, weights=w)
    assert abs(s[0] - patch_ncc_model_f64(f, m[0], 3, 1, mask=mask, weights=w)) < 5e-6
    # constant patches: sigma clamp, correlation 0, sim 1
    flat = np.zeros_like(f)
    assert abs(xo.patch_ncc(f, flat[None], o)[0] - 1.0) < 1e-6


def test_patch_grad_ncc_composition_and_combine(xo):
    f = _img(40, 44, seed=50)
    m = np.stack([_img(40, 44, seed=51), 0.6 * f + 0.3 * _img(40, 44, seed=52)])
    o = xo.patch_opts(radius=4)
    s = xo.patch_grad_ncc(f, m, o, gauss_width=5)
    fgx, fgy = xo.grad_imgs(f, 5)
    gx = np.stack([xo.grad_imgs(k, 5)[0] for k in m])
    gy = np.stack([xo.grad_imgs(k, 5)[1] for k in m])
    sx, sy = xo.patch_ncc(fgx, gx, o), xo.patch_ncc(fgy, gy, o)
    np.testing.assert_array_equal(s, (0.5 * (sx.astype(np.float64) + sy)).astype(f32))
    assert s[1] < s[0]
    np.testing.assert_allclose(xo.combine_mean(np.stack([sx, sy])), 0.5 * (sx + sy), rtol=1e-6)


# ---- pre-processing restated from ITK (parity unpinned: pinned here by independent implementations of the same algorithms) ----
def test_itk_discrete_gaussian_coefficients_are_the_discrete_gaussian(xo):
    """itk::GaussianOperator: e^-t I_n(t), normalised, summed until 1 - 0.01 -- against SciPy's exact Bessel functions."""
    import ctypes as C

    from scipy.special import ive

    lib = xo.lib()
    lib.xo_itk_gaussian_coeffs.restype = C.c_int
    for var, radius in ((2.0, 4), (1.0, 3), (4.0, 5), (16.0, 10), (0.25, 2)):
        k = (C.c_double * 160)()
        r = lib.xo_itk_gaussian_coeffs(C.c_double(var), C.c_double(0.01), 32, k)
        kk = np.array(k[: 2 * r + 1])
        ref = ive(np.arange(-r, r + 1), var)
        assert r == radius and abs(kk.sum() - 1.0) < 1e-14 and np.abs(kk - ref / ref.sum()).max() < 5e-9
        assert 2.0 * ive(np.arange(1, r), var).sum() + ive(0, var) < 0.99 <= 2.0 * ive(np.arange(1, r + 1), var).sum() + ive(0, var)


def test_downsample_image_matches_an_independent_bspline_implementation(xo):
    """DownsampleImage's resampling (cubic B-spline prefilter with mirror boundaries + 16-tap evaluation at i / factor)
    against SciPy's spline_filter / map_coordinates (mode 'mirror': the same published algorithm, another code base)."""
    from scipy import ndimage

    rng = np.random.default_rng(3)
    for shape, f in (((37, 53), 0.5), ((64, 64), 0.25), ((96, 80), 0.125), ((33, 29), 0.7), ((7, 5), 0.6)):
        img = rng.uniform(0, 100, shape).astype(f32)
        for sigma in (0.0, -1.0):
            got = xo.downsample_image(img, f, sigma)
            src = img if sigma == 0.0 else xo.itk_discrete_gaussian_2d(img, (0.5 / f) ** 2)
            c = ndimage.spline_filter(src.astype(np.float64), order=3, mode="mirror")
            yy, xx = np.meshgrid(np.arange(got.shape[0]) / f, np.arange(got.shape[1]) / f, indexing="ij")
            ref = ndimage.map_coordinates(c, [yy, xx], order=3, mode="mirror", prefilter=False)
            inside = (yy < shape[0] - 0.5) & (xx < shape[1] - 0.5)
            assert got.shape == (int(shape[0] * f + 0.5), int(shape[1] * f + 0.5))
            assert np.abs(got - ref)[inside].max() <= 1e-5 and np.all(got[~inside] == 0)
