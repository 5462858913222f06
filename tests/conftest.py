import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU is an error of invocation, not a pass: skip loudly
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device: GPU tests run under gpurun with -m gpu")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and load libs3g_b200.so; never falls back to anything."""
    from s3gaussian_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import splat_oracle
    splat_oracle.build()
    return splat_oracle
