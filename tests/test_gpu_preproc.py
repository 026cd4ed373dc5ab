"""GPU parity of the projection pre-processing: xrc_log_remap against the oracle's xo_log_remap, which
is pinned bit for bit to the reference's ImageIntensLogTransFilter::GenerateData lines (tests/test_oracle_ref_slice.py)."""
import numpy as np
import pytest

import xreg_b200

pytestmark = pytest.mark.gpu
f32 = np.float32


def _ulps(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize("shape", [(3, 5), (64, 80), (257, 301), (480, 480)])
def test_log_remap_matches_the_oracle(ctx, xo, shape):
    """I0 (the maximum of the smoothed image: the device runs the same restated ITK smoothing, double accumulation in tap
    order) and the value given to the non-positive pixels are bit-equal; the map itself is -log(x / I0) rounded once from a
    double logarithm, i.e. correctly rounded, while the CPU's std::log(float) (glibc logf, < 1 ulp) is not always:
    tolerance 1 ulp (1.2e-7 relative), seen on ~1 % of the pixels."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = (rng.uniform(0.0, 4000.0, shape) * (rng.random(shape) < 0.93)).astype(f32)
    img[rng.integers(0, shape[0]), rng.integers(0, shape[1])] = f32(5.0e-7)   # below eps
    for norm, use_max, i0 in ((False, True, 1.0), (True, True, 1.0), (False, False, 4096.0), (True, False, 2.5)):
        want, want_i0 = xo.log_remap(img, norm, use_max, i0)
        got, got_i0 = xreg_b200.log_remap(ctx, img, norm, use_max, i0)
        assert got_i0.tobytes() == want_i0.tobytes(), (norm, use_max)
        assert np.all(np.isfinite(got))
        d = _ulps(got, want)
        assert d.max() <= 1, d.max()
        assert np.count_nonzero(d) <= max(2, got.size // 20)
        low = img * (f32(1.0) / img.max() if norm else f32(1.0)) <= f32(1.0e-6)
        assert got[low].tobytes() == want[low].tobytes()      # the value of the smallest positive pixel


def test_log_remap_all_dark_and_errors(ctx, xo):
    img = np.zeros((8, 9), f32)
    want, _ = xo.log_remap(img, False, False, 1.0)
    got, _ = xreg_b200.log_remap(ctx, img, False, False, 1.0)
    assert np.array_equal(got, want) and np.all(np.isinf(got))     # min_pos stays 0 (:112): -log(0)
    with pytest.raises(xreg_b200.XregError):
        xreg_b200.log_remap(ctx, np.zeros((0, 4), f32))


@pytest.mark.parametrize("shape,factor", [((37, 53), 0.5), ((64, 64), 0.25), ((192, 160), 0.125), ((33, 29), 0.7),
                                          ((1, 40), 0.5), ((12, 9), 0.5), ((480, 480), 0.5)])
def test_downsample_image_is_the_oracle_bit_for_bit(ctx, xo, shape, factor):
    """xrc_downsample_image == xo_downsample_image (ITK's DownsampleImage chain restated: discrete Gaussian, cubic B-spline
    prefilter and evaluation): the same double arithmetic in the same order, uncontracted -> identical bytes; with the
    default smoothing, without, and with a given sigma; short lines take the prefilter's full-sum initialisation."""
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    img = rng.uniform(0.0, 3000.0, shape).astype(f32)
    for sigma in (-1.0, 0.0, 1.3):
        want = xo.downsample_image(img, factor, sigma)
        got = xreg_b200.downsample_image(ctx, img, factor, sigma)
        assert got.shape == want.shape and got.tobytes() == want.tobytes(), (sigma, float(np.abs(got - want).max()))


def test_downsample_proj_data(ctx, xo):
    """DownsampleProjData (lib/image/xregProjData.cpp:40-99): camera through DownsampleCameraModel, image through
    DownsampleImage, even dimensions forced by cropping from index (0, 0)."""
    from xreg_b200.geometry import CameraModel

    cam = CameraModel().setup(1020.0, 150, 190, 0.5, 0.5)
    img = np.random.default_rng(2).uniform(0, 100, (150, 190)).astype(f32)
    dimg, dcam = xreg_b200.downsample_proj_data(ctx, img, cam, 0.25, force_even_dims=True)
    assert dimg.shape == (dcam.num_det_rows, dcam.num_det_cols) == (38, 48)
    full = xo.downsample_image(img, 0.25)
    assert full.shape == (38, 48) and dimg.tobytes() == full[:38, :48].tobytes()
    dimg2, dcam2 = xreg_b200.downsample_proj_data(ctx, img[:, :150], CameraModel().setup(1020.0, 150, 150, 0.5, 0.5), 0.5 * 0.7)
    assert dimg2.shape == (dcam2.num_det_rows, dcam2.num_det_cols)
