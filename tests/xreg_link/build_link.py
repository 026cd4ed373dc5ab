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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
