"""the C++ adapter classes that plug libxreg_cuda.so into an xReg checkout (adapters/xreg/) derive
from the reference's real xreg::RayCaster / xreg::ImgSimMetric2D and include the reference's own headers, which need
Eigen / ITK / OpenCV / Boost.  Those libraries are absent here, so the adapters cannot be built into a binary; what CAN be
checked is that they type-check against the UNMODIFIED reference headers: tests/shim/ declares just enough of the
third-party API for `g++ -fsyntax-only`.  A wrong override signature, a missing include, a misspelt member of the
reference's base classes or an adapter class left abstract fails this test.

Runs only where the reference checkout exists (this container); skipped on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/lib"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

REF_DIRS = ["common", "ray_cast", "transforms", "regi/sim_metrics_2d", "itk", "opencv", "image", "hdf5", "regi",
            "basic_math", "file_formats"]
SOURCES = ["adapters/xreg/xregRayCastLineIntCUDA.cpp", "adapters/xreg/xregImgSimMetric2DCUDA.cpp",
           "tests/shim/adapters_instantiate.cpp"]


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.mark.parametrize("src", SOURCES)
def test_adapter_type_checks_against_the_reference_headers(src):
    inc = ["-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "adapters", "xreg")]
    for d in REF_DIRS:
        inc += ["-I", os.path.join(REF, d)]
    # C++11 like the reference (CMAKE_CXX_STANDARD 11); warnings in OUR sources are errors
    cmd = [_cxx(), "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra", "-Woverloaded-virtual", "-Wsuggest-override"] + inc + \
          [os.path.join(ROOT, src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ours = [ln for ln in r.stderr.splitlines() if "warning" in ln and "/root/reference" not in ln]
    assert not ours, "\n".join(ours)


def test_the_check_can_fail(tmp_path):
    """Guard against a vacuous check: an adapter with a wrong override signature must be rejected."""
    bad = tmp_path / "bad.cpp"
    bad.write_text('#include "xregRayCastInterface.h"\n'
                   'struct Bad : xreg::RayCaster { void compute(const int vol_idx) override; };\n')
    inc = ["-I", os.path.join(ROOT, "tests", "shim")]
    for d in REF_DIRS:
        inc += ["-I", os.path.join(REF, d)]
    r = subprocess.run([_cxx(), "-std=c++11", "-fsyntax-only"] + inc + [str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "override" in r.stderr
