"""s3gaussian_b200 - B200-native differentiable 4D Gaussian splatting hot path.

Drop-in for the render path of nnanhuang/S3Gaussian:

    s3gaussian_b200.diff_gaussian_rasterization   <- submodules/depth-diff-gaussian-rasterization
    s3gaussian_b200.gaussian_renderer.render       <- gaussian_renderer/__init__.py:23
    s3gaussian_b200.deformation                    <- scene/deformation.py + scene/hexplane.py

All compute goes through libs3g_b200.so (hand-written CUDA for sm_100a, C ABI in
include/s3g_b200.h).  There is no CPU fallback: importing is harmless, calling
any op without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_reference_module() -> None:
    """Make ``import diff_gaussian_rasterization`` resolve to this package's
    drop-in, so the reference's own ``gaussian_renderer/__init__.py:18`` import
    line works unchanged."""
    from . import diff_gaussian_rasterization as dgr
    sys.modules["diff_gaussian_rasterization"] = dgr
