import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from voxelraytrace20190722_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.bindings import Ref, build_ref
    if build_ref() is None:
        pytest.skip("oracle/_ref/libvrt_ref.so not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def gpu():
    """The product library; GPU tests FAIL (not skip) if it cannot run."""
    from voxelraytrace20190722_b200 import capi
    capi.load()
    assert capi.device_count() > 0, "no CUDA device: -m gpu tests must run on the GPU box"
    return capi
