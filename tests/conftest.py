import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def xo():
    """The CPU oracle (test infrastructure)."""
    from oracle import xreg_oracle

    xreg_oracle.build()
    return xreg_oracle


@pytest.fixture(scope="session")
def ctx():
    import xreg_b200

    c = xreg_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def small_scene():
    """64x64x48 anisotropic phantom + 96x80 detector that the volume fills."""
    from xreg_b200 import synth
    from xreg_b200.geometry import CameraModel

    vol = synth.make_volume(64, 64, 48, spacing=(0.9, 1.1, 1.3))
    cam = CameraModel().setup(400.0, 80, 96, 1.6, 1.5)
    nominal = synth.nominal_pose(vol, src_to_iso=250.0)
    return vol, cam, nominal
