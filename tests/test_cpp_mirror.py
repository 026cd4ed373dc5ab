"""The C++ host mirror (include/xreg_cuda.hpp: RayCaster / ImgSimMetric2D with the reference's names and call-order
contract over the C ABI) driven by a C++ test program, tests/cpp/host_mirror_test.cpp, which compares with the CPU
oracle.  The oracle is linked into the TEST binary only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build(tmp_path, xo):
    from xreg_b200 import _lib

    _lib.load()  # builds libxreg_cuda.so on a fresh checkout
    xo.build()
    exe = str(tmp_path / "host_mirror_test")
    libdir, odir = os.path.join(ROOT, "xreg_b200"), os.path.join(ROOT, "oracle")
    cmd = [_cxx(), "-std=c++14", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
           "-I", os.path.join(ROOT, "include"), "-I", odir, os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
           "-o", exe, "-L", libdir, "-lxreg_cuda", "-L", odir, "-lxreg_oracle",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_cxx11(tmp_path):
    """The reference builds as C++11 (CMakeLists.txt: CMAKE_CXX_STANDARD 11): the mirror must too, pedantically."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "xreg_cuda.hpp"\nint main() { xreg_b200::CameraModel c; c.setup(100.f, 8, 8, 1.f, 1.f); '
                   'return c.num_det_rows == 8 ? 0 : 1; }\n')
    from xreg_b200 import _lib

    _lib.load()
    libdir = os.path.join(ROOT, "xreg_b200")
    exe = tmp_path / "t"
    subprocess.run([_cxx(), "-std=c++11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lxreg_cuda", "-Wl,-rpath," + libdir], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_cpp_mirror_host_logic(tmp_path, xo):
    """CameraModel set-up, transform algebra, patch grid / weights bit-equal to the oracle; exception types."""
    exe = _build(tmp_path, xo)
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_mirror_test (no gpu): ok" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_parity_on_gpu(tmp_path, xo):
    """Whole path through the C++ classes: DRR (bit-exact masks / sample counts, <= 1e-4), five metrics (<= 1e-5),
    multi-view distribution + CombineMean, one-call objective, store methods, error behaviour."""
    exe = _build(tmp_path, xo)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_mirror_test: ok" in r.stdout
