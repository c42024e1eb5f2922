import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(autouse=True)
def _production_factor_kernel(monkeypatch):
    """Small batches would get the one-item-per-warp factor kernel by the library's batch-size rule; the tests pin the
    two-items-per-warp kernel the large batches of the bench run, unless a test (or the environment) chooses itself."""
    import os

    if "LVIO2D_FACTOR_PAIRED" not in os.environ:
        monkeypatch.setenv("LVIO2D_FACTOR_PAIRED", "1")
    yield


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def params():
    import lvio2d_b200 as L

    return L.corridor_params()


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def pytest_collection_modifyitems(config, items):
    """GPU tests call the in-tree CUDA library; rebuild it first when a source is newer (nvcc is in the image)."""
    if not any("gpu" in item.keywords for item in items):
        return
    csrc = os.path.join(ROOT, "2dliw-slam_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))] + [os.path.join(ROOT, "include", "lvio2d.h")]
    if _stale(os.path.join(csrc, "liblvio2d.so"), srcs):
        import subprocess

        subprocess.check_call(["bash", os.path.join(csrc, "build.sh")])
