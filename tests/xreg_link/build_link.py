#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Builds oracle/_ref/xreg_adapter_driver: the xReg adapter classes
(adapters/xreg/*.cpp) compiled and LINKED together with the reference's own base-class sources, taken where they lie
under /root/reference:

  whole files   lib/common/xregExceptionUtils.cpp, xregAssert.cpp, xregSampleUtils.cpp
                lib/ray_cast/xregRayCastInterface.cpp, xregRayCastSyncBuf.cpp
                lib/regi/sim_metrics_2d/xregImgSimMetric2D.cpp, xregImgSimMetric2DPatchCommon.cpp,
                                        xregImgSimMetric2DCombine.cpp
  by anchor     lib/transforms/xregPerspectiveXform.cpp: FocalLenFromIntrins, MakeNaiveIntrins, both CameraModel::setup
                overloads the tests use (the rest of that file needs Eigen's SVD / QR);
                lib/transforms/xregRigidUtils.cpp: SE3Inv

against the functional Eigen / ITK / OpenCV / Boost stand-ins of tests/xreg_link/third_party (declaration-only ones from
tests/shim for everything that is only mentioned), -DXREG_NO_TBB (the reference's own serial switch,
lib/common/xregTBBUtils.h:36-43), plus tests/xreg_link/driver.cpp, and libxreg_cuda.so.  No reference source is copied
into the repository: the generated slice is deleted after compiling, objects live in a temporary directory, the only
output is the binary under oracle/_ref/ (git-ignored, shipped to the GPU box with the snapshot).
A no-op where /root/reference is absent (the GPU box uses the shipped binary)."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_pin.build_ref_slice import REF, _cut_function, _lines  # noqa: E402

OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(OUT_DIR, "xreg_adapter_driver")
REF_DIRS = ["common", "ray_cast", "transforms", "regi/sim_metrics_2d", "itk", "opencv", "image", "hdf5", "regi",
            "basic_math", "file_formats", "spatial"]
WHOLE = ["lib/common/xregExceptionUtils.cpp", "lib/common/xregAssert.cpp", "lib/common/xregSampleUtils.cpp",
         "lib/ray_cast/xregRayCastInterface.cpp", "lib/ray_cast/xregRayCastSyncBuf.cpp",
         "lib/regi/sim_metrics_2d/xregImgSimMetric2D.cpp", "lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchCommon.cpp",
         "lib/regi/sim_metrics_2d/xregImgSimMetric2DCombine.cpp"]
OURS = ["adapters/xreg/xregRayCastLineIntCUDA.cpp", "adapters/xreg/xregImgSimMetric2DCUDA.cpp", "tests/xreg_link/driver.cpp"]


def _cam_slice():
    out = ['#include "xregPerspectiveXform.h"', '#include "xregRigidUtils.h"', '#include "xregAssert.h"', "#include <cmath>", ""]
    ln = _lines("lib/transforms/xregRigidUtils.cpp")
    s, e = _cut_function(ln, r"^xreg::Mat4x4 xreg::SE3Inv\(const Mat4x4& T\)")
    out += ln[s:e + 1] + [""]
    ln = _lines("lib/transforms/xregPerspectiveXform.cpp")
    for rx in (r"^xreg::CoordScalar xreg::FocalLenFromIntrins\(", r"^xreg::Mat3x3 xreg::MakeNaiveIntrins\(",
               r"^void xreg::CameraModel::setup\(const CoordScalar focal_len_arg",
               r"^void xreg::CameraModel::setup\(const Mat3x3& intrins_mat, const Mat4x4& extrins_mat"):
        s, e = _cut_function(ln, rx)
        out += ln[s:e + 1] + [""]
    return "\n".join(out)


def includes():
    inc = ["-I", os.path.join(HERE, "third_party"), "-I", os.path.join(ROOT, "tests", "shim"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "adapters", "xreg")]
    for d in REF_DIRS:
        inc += ["-I", os.path.join(REF, "lib", d)]
    return inc


def build(force: bool = False) -> str:
    if not os.path.isdir(REF):
        return BIN
    srcs = [os.path.join(REF, w) for w in WHOLE] + [os.path.join(ROOT, o) for o in OURS] + [__file__,
            os.path.join(ROOT, "xreg_b200", "libxreg_cuda.so")]
    hdrs = []
    for d, _, fs in os.walk(os.path.join(HERE, "third_party")):
        hdrs += [os.path.join(d, f) for f in fs]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(s) for s in srcs + hdrs if os.path.exists(s)):
        return BIN
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="xreg_link_")
    try:
        cxx = ["g++", "-std=c++11", "-O1", "-ffp-contract=off", "-DXREG_NO_TBB", "-include", "cmath", "-include", "set"] + includes()
        objs = []
        slice_cpp = os.path.join(tmp, "cam_slice.cpp")
        with open(slice_cpp, "w") as f:
            f.write(_cam_slice())
        for i, src in enumerate([os.path.join(REF, w) for w in WHOLE] + [slice_cpp] + [os.path.join(ROOT, o) for o in OURS]):
            obj = os.path.join(tmp, "o%d.o" % i)
            subprocess.run(cxx + ["-c", src, "-o", obj], check=True)
            objs.append(obj)
        os.remove(slice_cpp)
        subprocess.run(["g++", "-o", BIN] + objs + ["-L", os.path.join(ROOT, "xreg_b200"), "-lxreg_cuda",
                                                    "-Wl,-rpath,$ORIGIN/../../xreg_b200", "-Wl,--allow-shlib-undefined"], check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return BIN


# ---- the reference's SE(3) magnitude penalty (Regi2D3DPenaltyFnSE3Mag + FoldNormDist) as a shared library, to pin
# xrc_se3_mag_penalty: whole files lib/basic_math/xregFoldNormDist.cpp, xregDistInterface.cpp,
# lib/regi/penalty_fns_2d_3d/xregRegi2D3DPenaltyFn.cpp, xregRegi2D3DPenaltyFnSE3Mag.cpp; LogSO3ToPt
# (xregRotUtils.cpp:107-126) and ComputeRotAngTransMag (xregRigidUtils.cpp:246-251) by anchor
PEN_LIB = os.path.join(OUT_DIR, "libxreg_refpenalty.so")
PEN_WHOLE = ["lib/basic_math/xregFoldNormDist.cpp", "lib/basic_math/xregDistInterface.cpp",
             "lib/regi/penalty_fns_2d_3d/xregRegi2D3DPenaltyFn.cpp", "lib/regi/penalty_fns_2d_3d/xregRegi2D3DPenaltyFnSE3Mag.cpp",
             "lib/common/xregExceptionUtils.cpp", "lib/common/xregAssert.cpp"]
PEN_WRAPPER = r'''
#include <memory>
#include "xregRegi2D3DPenaltyFnSE3Mag.h"
#include "xregFoldNormDist.h"
#include "xregPerspectiveXform.h"

static xreg::FrameTransform from12(const float* a)
{
  xreg::FrameTransform t;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c)
      t.matrix()(r, c) = a[4 * r + c];
  return t;
}

extern "C" int xref_se3_mag_penalty(float rot_m, float rot_s, float trans_m, float trans_s, int inter_wrt_vol,
                                    const float* inter12, const float* init12, unsigned n, const float* cams12, float* out)
{
  xreg::Regi2D3DPenaltyFnSE3Mag pen;
  pen.rot_pdfs_per_obj = {std::make_shared<xreg::FoldNormDist>(rot_m, rot_s)};
  pen.trans_pdfs_per_obj = {std::make_shared<xreg::FoldNormDist>(trans_m, trans_s)};
  xreg::Regi2D3DPenaltyFn::ListOfFrameTransformLists cams(1);
  for (unsigned p = 0; p < n; ++p)
    cams[0].push_back(from12(cams12 + 12 * p));
  pen.compute(cams, n, xreg::Regi2D3DPenaltyFn::CamList(), xreg::Regi2D3DPenaltyFn::CamAssocList(), {inter_wrt_vol != 0},
              {from12(inter12)}, {from12(init12)}, nullptr);
  for (unsigned p = 0; p < n; ++p)
    out[p] = pen.reg_vals()[p];
  return 0;
}
'''


def _pen_slice():
    out = ['#include "xregRotUtils.h"', '#include "xregRigidUtils.h"', "#include <cmath>", "#include <tuple>", ""]
    ln = _lines("lib/transforms/xregRotUtils.cpp")
    s, e = _cut_function(ln, r"^xreg::Pt3 xreg::LogSO3ToPt\(const Mat3x3& R\)")
    out += ln[s:e + 1] + [""]
    ln = _lines("lib/transforms/xregRigidUtils.cpp")
    s, e = _cut_function(ln, r"^xreg::ComputeRotAngTransMag\(const FrameTransform& xform\)")
    out += ln[s - 1:e + 1] + [""]     # the return type sits on the line before
    return "\n".join(out)


def build_penalty(force: bool = False) -> str:
    if not os.path.isdir(REF):
        return PEN_LIB
    if not force and os.path.exists(PEN_LIB) and os.path.getmtime(PEN_LIB) >= os.path.getmtime(__file__):
        return PEN_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="xreg_pen_")
    try:
        inc = includes() + ["-I", os.path.join(REF, "lib", "regi", "penalty_fns_2d_3d")]
        cxx = ["g++", "-std=c++11", "-O1", "-ffp-contract=off", "-fPIC", "-DXREG_NO_TBB", "-include", "cmath"] + inc
        srcs = [os.path.join(REF, w) for w in PEN_WHOLE]
        for name, text in (("pen_slice.cpp", _pen_slice()), ("pen_wrapper.cpp", PEN_WRAPPER)):
            path = os.path.join(tmp, name)
            with open(path, "w") as f:
                f.write(text)
            srcs.append(path)
        objs = []
        for i, src in enumerate(srcs):
            obj = os.path.join(tmp, "p%d.o" % i)
            subprocess.run(cxx + ["-c", src, "-o", obj], check=True)
            objs.append(obj)
        subprocess.run(["g++", "-shared", "-o", PEN_LIB] + objs, check=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return PEN_LIB


if __name__ == "__main__":
    print(build_penalty(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv))
