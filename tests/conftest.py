import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """a fresh checkout has no built library (it is git-ignored): build it once, like the driver's build() step"""
    lib = os.path.join(ROOT, "matrixproductbp.jl_b200", "libmpbp_b200.so")
    if not os.path.exists(lib):
        import shutil
        if shutil.which("nvcc"):
            import __graft_entry__ as G
            G.build()
