"""pytest configuration: markers, repo-root imports, oracle build."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes binding of the CPU oracle (oracle/libzk_oracle.so), built on demand."""
    so = os.path.join(ROOT, "oracle", "libzk_oracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    from tests.oracle_binding import Oracle

    return Oracle(so)


def _have_gpu():
    try:
        import zlibk_b200

        return zlibk_b200.library().zk_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C ABI.  GPU tests fail loudly (not skip) when the library is missing."""
    import zlibk_b200

    return zlibk_b200.Engine(0)


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
