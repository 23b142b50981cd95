import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def golden_cases(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))


def pytest_sessionstart(session):
    # keep the in-tree library in sync with its sources during development (mtime based)
    try:
        from feonet_navier_stokes_b200 import build as _b

        if _b.needs_build():
            _b.build_library()
    except Exception as exc:  # pragma: no cover
        print(f"[conftest] could not (re)build libfeonet_b200.so: {exc}")
