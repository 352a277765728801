import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the in-tree libraries (host .so, CUDA .so by cross-compilation, oracle, shim) once."""
    from rendering_b200 import build
    import subprocess
    build.build_all()
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "librtb_oracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "restate"], check=True, capture_output=True)
    yield
