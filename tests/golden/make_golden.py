"""Generates tests/golden/small_scene.npz with the CPU oracle (run here, in the build container):

    python tests/golden/make_golden.py

The reference itself cannot be built or imported in this image (C++ with ITK / Eigen / OpenCV /
TBB, none present), so these vectors come from the oracle restatement, not from xReg; they pin
the oracle against drift across compilers / machines and give the GPU tests a second anchor.
Inputs are regenerated from seeds by xreg_b200.synth; only poses and outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import xreg_oracle as xo  # noqa: E402
from xreg_b200 import synth  # noqa: E402
from xreg_b200.geometry import CameraModel, to12  # noqa: E402


def scene():
    vol = synth.make_volume(48, 40, 36, spacing=(1.0, 1.2, 1.4), seed=123)
    cam = CameraModel().setup(420.0, 40, 48, 3.0, 2.8)
    nominal = synth.nominal_pose(vol, src_to_iso=260.0)
    poses = synth.pose_population(vol, nominal, 4, seed=321, sigma=(8, 8, 8, 5, 5, 8))
    return vol, cam, poses


def main():
    vol, cam, poses = scene()
    cams = [xo.cam_struct(cam)]
    drr, mask, steps, S = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), want_info=True)
    fixed = synth.add_noise(drr[0], seed=77)
    cmask = synth.circular_mask(*fixed.shape)
    o = xo.patch_opts(radius=4)
    out = dict(
        poses=poses, drr=drr, mask=np.packbits(mask), steps=steps.astype(np.uint16), S=np.uint64(S), fixed=fixed,
        ncc=xo.ncc(fixed, drr), grad_ncc=xo.grad_ncc(fixed, drr), patch_ncc=xo.patch_ncc(fixed, drr, o),
        patch_grad_ncc=xo.patch_grad_ncc(fixed, drr, o),
        patch_grad_ncc_masked=xo.patch_grad_ncc(fixed, drr, o, mask=cmask, weights=xo.patch_weights(40, 48, o, mask=cmask)),
        grad_ncc_masked=xo.grad_ncc(fixed, drr, mask=cmask),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_scene.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; S =", S)


def main_round2():
    """tests/golden/round2.npz: the round-2 additions on the same scene (nearest-neighbour DRR, depth images, log remap and
    down-sampling of the fixed image), by the oracle."""
    vol, cam, poses = scene()
    cams = [xo.cam_struct(cam)]
    drr = xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses))
    fixed = synth.add_noise(drr[0], seed=77)
    intens = (4000.0 * np.exp(-fixed)).astype(np.float32)     # an intensity image, as a detector would deliver it
    intens[3, 7] = 0.0
    vmax = float(vol.data.max())
    log_img, log_i0 = xo.log_remap(intens)
    out = dict(
        drr_nn=xo.drr(vol.data, vol.idx_to_phys(), cams, to12(poses), interp=1),
        depth=xo.depth(vol.data, vol.idx_to_phys(), cams, to12(poses), thresh=0.5 * vmax, n_backtrack=4),
        depth_nn=xo.depth(vol.data, vol.idx_to_phys(), cams, to12(poses), interp=1, thresh=0.3 * vmax, n_backtrack=0),
        depth_thresh=np.float32(0.5 * vmax), depth_nn_thresh=np.float32(0.3 * vmax),
        intens=intens, log_remap=log_img, log_i0=log_i0,
        down_half=xo.downsample_image(intens, 0.5), down_quarter_nosmooth=xo.downsample_image(intens, 0.25, 0.0),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "round2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main_round2()
