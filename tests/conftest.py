import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build them once, the way
    __graft_entry__.build() does, instead of failing every test on import.  The PRODUCT still
    fails loudly when its library is missing; this is only the test session helping itself."""
    import shutil
    import subprocess
    so = os.path.join(ROOT, "manatee_b200", "libmanatee_gpu.so")
    if not os.path.exists(so) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "manatee_b200", "csrc")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def native():
    """libmanatee_gpu.so must be built and loadable; GPU tests fail loudly otherwise."""
    from manatee_b200 import _native
    return _native.lib()
