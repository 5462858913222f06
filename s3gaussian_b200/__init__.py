"""s3gaussian_b200 - B200-native differentiable 4D Gaussian splatting hot path.

Drop-in for the render path of nnanhuang/S3Gaussian:

    s3gaussian_b200.diff_gaussian_rasterization   <- submodules/depth-diff-gaussian-rasterization
    s3gaussian_b200.gaussian_renderer.render       <- gaussian_renderer/__init__.py:23
    s3gaussian_b200.deformation                    <- scene/deformation.py + scene/hexplane.py
    s3gaussian_b200.gaussian_model.GaussianModel   <- scene/gaussian_model.py (training state, densify / prune, .ply)
    s3gaussian_b200.optim / losses / regulation    <- torch.optim.Adam call site, utils/loss_utils.py, scene/regulation.py
    s3gaussian_b200.simple_knn.distCUDA2           <- submodules/simple-knn

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
    # `from simple_knn._C import distCUDA2` (scene/gaussian_model.py:24)
    import types
    from . import simple_knn as knn
    pkg = types.ModuleType("simple_knn")
    pkg.__path__ = []
    ext = types.ModuleType("simple_knn._C")
    ext.distCUDA2 = knn.distCUDA2
    pkg._C = ext
    sys.modules["simple_knn"] = pkg
    sys.modules["simple_knn._C"] = ext
