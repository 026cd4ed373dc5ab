"""C-ABI boundary checks that need no GPU: the shared library loads, exports every
symbol include/xreg_cuda.h declares, the POD layouts agree, and compute entry points
fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "xreg_cuda.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xrc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from xreg_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libxreg_cuda.so does not export %s" % n
    # and the Python binding declares a signature for each of them
    assert set(names) == set(_lib.SIGNATURES.keys())


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "xreg_cuda.h"\nint main(void){ xrc_cam c; (void)c; return sizeof(xrc_cam) == 112 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_cam_struct_layouts_agree(xo):
    from xreg_b200 import _lib

    assert C.sizeof(_lib.XrcCam) == C.sizeof(xo.XoCam) == 112
    for (n1, t1), (n2, t2) in zip(_lib.XrcCam._fields_, xo.XoCam._fields_):
        assert n1 == n2 and C.sizeof(t1) == C.sizeof(t2)
        assert getattr(_lib.XrcCam, n1).offset == getattr(xo.XoCam, n2).offset


def test_version_and_launch_counter():
    from xreg_b200 import _lib

    lib = _lib.load()
    assert lib.xrc_version() == 100
    assert lib.xrc_launch_count() >= 0


def test_null_arguments_are_rejected_without_a_device():
    from xreg_b200 import _lib

    lib = _lib.load()
    assert lib.xrc_ctx_create(0, None) == _lib.XRC_ERR_INVALID
    assert b"null" in lib.xrc_last_error()
    assert lib.xrc_rc_create(None, None) == _lib.XRC_ERR_INVALID
    assert lib.xrc_sm_compute(None) == _lib.XRC_ERR_INVALID
    assert lib.xrc_rc_compute(None, 0) == _lib.XRC_ERR_INVALID


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import xreg_b200

    with pytest.raises(xreg_b200.XregCudaError) as e:
        xreg_b200.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "xreg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "xreg_oracle" not in txt and "libxreg_oracle" not in txt, f
    # the shared library does not link the oracle either
    out = subprocess.run(["ldd", os.path.join(pkg, "libxreg_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def _build_c_client(tmp_path):
    """tests/c_abi/abi_smoke.c: a plain-C99 client of include/xreg_cuda.h linked against libxreg_cuda.so."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(root, "xreg_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "c_abi", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lxreg_cuda", "-lm",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_valid_c99_and_c_client_links(tmp_path):
    """The boundary is a C ABI: the header must compile as strict C99 and a C program must link and run the
    entry points that need no device (version, exp map, error convention)."""
    import subprocess

    from xreg_b200 import _lib

    _lib.load()  # builds the library on a fresh checkout
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


@pytest.mark.gpu
def test_c_client_end_to_end_on_gpu(tmp_path):
    """The same C program, whole path: analytic line integral of a constant volume, NCC through xrc_obj_fn."""
    import subprocess

    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke: ok" in r.stdout
