"""pytest configuration: registers the ``gpu`` marker and puts the repo root /
oracle on sys.path.  ``-m "not gpu"`` runs on the CPU-only dev container,
``-m gpu`` on a B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared object is git-ignored: build it when a fresh checkout runs the suite before __graft_entry__.build()
    lib = os.path.join(ROOT, "brickbrain_rec_engine_b200", "libbrickscore.so")
    if not os.path.exists(lib):
        from brickbrain_rec_engine_b200 import build as _b
        _b.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
