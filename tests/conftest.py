import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs (a dead peer, a wedged stream) must end the run with a failure, not sit there until the box is reclaimed: every
    gpu-marked test gets a 10-minute limit (the whole suite takes about half a minute).  method=thread: the watchdog ends the process even
    when the main thread is blocked inside a CUDA call, where a signal handler would never get to run."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))


@pytest.fixture(scope="session")
def tmm():
    """The product package.  Loading fails loudly when the CUDA extension is not built."""
    import tiled_mm_b200
    tiled_mm_b200.load_library()
    return tiled_mm_b200


@pytest.fixture(scope="session")
def oracle():
    import _util
    return _util.Oracle()


@pytest.fixture(scope="session")
def gpu_tmm(tmm):
    """Product package on a box that must have a GPU: no device => the test FAILS (no silent CPU path)."""
    n = tmm.device_count()
    assert n >= 1, "pytest -m gpu needs a CUDA device; tiled_mm_b200 has no CPU fallback"
    return tmm
