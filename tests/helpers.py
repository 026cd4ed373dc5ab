"""Shared test helpers: an independent numpy float64 model of the DRR and metrics
(second implementation of SURVEY Appendix A, used to cross-check the C oracle)."""
import numpy as np


def drr_model_f64(vol, idx_to_phys, cam, pose44, step=1.0):
    """Line integral DRR in float64, vectorised over pixels.  Returns (img, hit_mask)."""
    nz, ny, nx = vol.shape
    A = np.eye(4)
    A[:3, :] = np.asarray(idx_to_phys, dtype=np.float64).reshape(3, 4)
    X = np.linalg.inv(A) @ np.asarray(pose44, dtype=np.float64)
    rows, cols = cam.num_det_rows, cam.num_det_cols
    cc, rr = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    det_z = (-1.0 if cam.coord_frame_type == 1 else 1.0) * float(cam.focal_len)
    ind = np.stack([cc * det_z, rr * det_z, np.full_like(cc, det_z)], axis=-1)
    p3 = ind @ np.asarray(cam.intrins_inv, dtype=np.float64).T
    if cam.coord_frame_type == 2:
        p3[..., 2] -= float(cam.focal_len)
    E = np.asarray(cam.extrins_inv, dtype=np.float64)
    det = p3 @ E[:3, :3].T + E[:3, 3]
    ph = np.asarray(cam.pinhole_pt, dtype=np.float64)
    p = X[:3, :3] @ ph + X[:3, 3]
    d = det @ X[:3, :3].T + X[:3, 3] - p
    mx = np.array([nx - 1, ny - 1, nz - 1], dtype=np.float64)
    t0 = np.zeros((rows, cols))
    t1 = np.ones((rows, cols))
    hit = np.ones((rows, cols), dtype=bool)
    for k in range(3):
        par = np.abs(d[..., k]) <= 1e-8
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d[..., k]
            a = (0.0 - p[k]) * inv
            b = (mx[k] - p[k]) * inv
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        t0 = np.where(par, t0, np.maximum(t0, lo))
        t1 = np.where(par, t1, np.minimum(t1, hi))
        hit &= np.where(par, (p[k] >= 0) & (p[k] <= mx[k]), True)
    hit &= (t0 <= t1) & ((t1 - t0) > 2e-3)
    t0 = t0 + 1e-3
    t1 = t1 - 1e-3
    L = np.linalg.norm(d, axis=-1)
    u = det - ph
    u = u / np.linalg.norm(u, axis=-1, keepdims=True) * step
    step_len = np.linalg.norm(u @ X[:3, :3].T, axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        nsteps = np.where(hit, np.floor((t1 - t0) * L / step_len), -1).astype(np.int64)
    stepv = d * (step_len / L)[..., None]
    start = p + t0[..., None] * d
    img = np.zeros((rows, cols))
    v = vol.astype(np.float64)
    for s in range(int(nsteps.max()) + 1 if hit.any() else 0):
        act = hit & (s <= nsteps)
        x = start + s * stepv
        x = np.clip(x, 0, mx)
        i0 = np.minimum(np.floor(x).astype(np.int64), (mx - 0).astype(np.int64))
        w = x - i0
        i1 = np.minimum(i0 + 1, mx.astype(np.int64))

        def g(ix, iy, iz):
            return v[iz, iy, ix]

        c00 = g(i0[..., 0], i0[..., 1], i0[..., 2]) * (1 - w[..., 0]) + g(i1[..., 0], i0[..., 1], i0[..., 2]) * w[..., 0]
        c10 = g(i0[..., 0], i1[..., 1], i0[..., 2]) * (1 - w[..., 0]) + g(i1[..., 0], i1[..., 1], i0[..., 2]) * w[..., 0]
        c01 = g(i0[..., 0], i0[..., 1], i1[..., 2]) * (1 - w[..., 0]) + g(i1[..., 0], i0[..., 1], i1[..., 2]) * w[..., 0]
        c11 = g(i0[..., 0], i1[..., 1], i1[..., 2]) * (1 - w[..., 0]) + g(i1[..., 0], i1[..., 1], i1[..., 2]) * w[..., 0]
        c0 = c00 * (1 - w[..., 1]) + c10 * w[..., 1]
        c1 = c01 * (1 - w[..., 1]) + c11 * w[..., 1]
        val = c0 * (1 - w[..., 2]) + c1 * w[..., 2]
        img += np.where(act, val, 0.0)
    return img * step, hit, nsteps


def ncc_model_f64(fixed, mov, mask=None):
    f = fixed.astype(np.float64).ravel()
    m = mov.astype(np.float64).ravel()
    if mask is not None:
        sel = mask.ravel() != 0
        f, m = f[sel], m[sel]
    n = f.size
    f0, m0 = f - f.mean(), m - m.mean()
    sf = max(1e-6, np.sqrt((f0 ** 2).sum() / (n - 1)))
    sm = max(1e-6, np.sqrt((m0 ** 2).sum() / (n - 1)))
    return 0.5 * (1.0 - (f0 @ m0) / (n * sf * sm))


def patch_ncc_model_f64(fixed, mov, radius, stride=1, mask=None, weights=None, weight_sims=True, mean=False):
    """Direct O(P d^2) float64 model of ImgSimMetric2DPatchNCCCPU with default mask handling."""
    rows, cols = fixed.shape
    f = fixed.astype(np.float64)
    m = mov.astype(np.float64)
    d = 2 * radius + 1
    n = d * d
    sims = []
    for cr in range(radius, rows - radius, stride):
        for cc in range(radius, cols - radius, stride):
            fp = f[cr - radius:cr + radius + 1, cc - radius:cc + radius + 1]
            mp = m[cr - radius:cr + radius + 1, cc - radius:cc + radius + 1]
            sf = max(1e-6, np.sqrt(((fp - fp.mean()) ** 2).sum() / (n - 1)))
            sm = max(1e-6, np.sqrt(((mp - mp.mean()) ** 2).sum() / (n - 1)))
            prod = ((mp - mp.mean()) / sm) * ((fp - fp.mean()) / (sf * n))
            if mask is not None:
                prod = prod * (mask[cr - radius:cr + radius + 1, cc - radius:cc + radius + 1] != 0)
            sims.append(1.0 - prod.sum())
    sims = np.array(sims)
    w = np.ones_like(sims) if weights is None else weights.astype(np.float64)
    if weight_sims:
        use = np.abs(w) > 1e-6
        tot = (w * sims * use).sum()
    else:
        tot = sims.sum()
    if mean:
        return tot / sims.size
    if weight_sims:
        return tot / w.sum()
    return tot
