"""``distCUDA2`` of the reference's simple-knn submodule (submodules/simple-knn/simple_knn.cu:185-221,
``from simple_knn._C import distCUDA2`` at scene/gaussian_model.py:24,153): mean squared distance of every
point to its three nearest neighbours, used to initialise the Gaussian scales.  Exact, same arithmetic,
our own search structure (csrc/knn.cuh).  CUDA only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("distCUDA2: CUDA tensor expected (there is no CPU path)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError(f"distCUDA2: points must be [P,3], got {tuple(points.shape)}")
    pts = points.detach().float().contiguous()
    P = pts.shape[0]
    out = torch.full((P,), 0.0, device=pts.device, dtype=torch.float32)      # spatial.cu:21
    if P == 0:
        return out
    lib = _lib.load()
    ws = torch.empty(lib.s3g_knn_workspace_bytes(P), dtype=torch.uint8, device=pts.device)
    _lib.check(lib.s3g_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), ws.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s3g_knn_mean_dist2")
    return out
