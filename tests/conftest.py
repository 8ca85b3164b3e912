import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ref():
    """the compiled reference (golden build: SIMD=0, shared fp32 math)"""
    from oracle.pyref import RefLib, have_ref
    if not have_ref("scalar"):
        pytest.skip("oracle/_ref/libvgref.so not built (make -C oracle ref needs /root/reference)")
    return RefLib("scalar")


@pytest.fixture(scope="session")
def gpu_batch():
    import vg_renderer_b200 as vgb
    if vgb.device_count() < 1:
        pytest.fail("gpu-marked test on a box without a CUDA device")
    return vgb.Batch(0)
