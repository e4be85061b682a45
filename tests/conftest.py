import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build anything that is missing (libvrt.so / libvrt_host.so / oracle) — nvcc cross-compiles without a GPU
    need = [os.path.join(ROOT, "zig_vulkan_b200", "libvrt.so"), os.path.join(ROOT, "zig_vulkan_b200", "libvrt_host.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", ROOT, "all"], check=True, capture_output=True)


POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))


@pytest.fixture(scope="session")
def materials():
    import zig_vulkan_b200 as zv

    return zv.terrain_materials()


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
