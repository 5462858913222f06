"""HexPlane regularisers of the fine stage as one forward and one backward launch over all planes
(SURVEY section 8f row f-3).

The reference evaluates per fine iteration (train.py:414-417):

    tv_loss = gaussians.compute_regulation(hyper.time_smoothness_weight, hyper.l1_time_planes, hyper.plane_tv_weight)

= plane_tv_weight * sum(smoothness of planes 0,1,3) + time_smoothness_weight * sum(smoothness of planes
2,4,5) + l1_time_planes * sum(mean|1 - plane| of planes 2,4,5) over every resolution level
(scene/gaussian_model.py:710-749, compute_plane_smoothness scene/regulation.py:22-28) - about ten torch
launches and five full-size temporaries per plane, forward and again backward.  Here: one read pass
forward, one read + one write pass backward, for all 24 planes together.  CUDA fp32 only; no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

SPATIAL_PLANES = (0, 1, 3)      # (x,y), (x,z), (y,z)
TIME_PLANES = (2, 4, 5)         # (x,t), (y,t), (z,t)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _physical(p: torch.Tensor):
    """-> (H, W, C) of a [1,C,H,W] plane whose memory is [H][W][C] (channels_last)."""
    if p.dim() != 4 or p.shape[0] != 1:
        raise RuntimeError(f"plane_regulation: expected [1,C,H,W] planes, got {tuple(p.shape)}")
    if not p.is_cuda or p.dtype != torch.float32:
        raise RuntimeError("plane_regulation: CUDA float32 planes only (there is no CPU path)")
    if not p.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("plane_regulation: planes must be in torch.channels_last memory format "
                           "(s3gaussian_b200.deformation.HexPlaneField stores them that way)")
    return p.shape[2], p.shape[3], p.shape[1]


class _PlaneRegulation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeffs, *planes):
        lib = _lib.load()
        descs = []
        for (ws, wl), p in zip(coeffs, planes):
            H, W, Cc = _physical(p)
            descs.append(_lib.PlaneDesc(p.data_ptr(), None, H, W, Cc, float(ws), float(wl)))
        arr = (_lib.PlaneDesc * len(descs))(*descs)
        dev = planes[0].device
        ws = torch.empty(lib.s3g_plane_reg_workspace_bytes(len(descs), arr), dtype=torch.uint8, device=dev)
        total = torch.empty(1, dtype=torch.float64, device=dev)
        _lib.check(lib.s3g_plane_reg_forward(len(descs), arr, total.data_ptr(), ws.data_ptr(), _stream()),
                   "s3g_plane_reg_forward")
        ctx.coeffs = coeffs
        ctx.save_for_backward(*planes)
        return total[0].float()

    @staticmethod
    def backward(ctx, g):
        planes = ctx.saved_tensors
        lib = _lib.load()
        grads = [torch.empty_like(p, memory_format=torch.preserve_format) for p in planes]
        descs = []
        for (ws, wl), p, gr in zip(ctx.coeffs, planes, grads):
            H, W, Cc = _physical(p)
            descs.append(_lib.PlaneDesc(p.data_ptr(), gr.data_ptr(), H, W, Cc, float(ws), float(wl)))
        arr = (_lib.PlaneDesc * len(descs))(*descs)
        gs = g.reshape(1).to(torch.float32).contiguous()
        _lib.check(lib.s3g_plane_reg_backward(len(descs), arr, gs.data_ptr(), _stream()), "s3g_plane_reg_backward")
        return (None, *grads)


def compute_regulation(multi_res_grids, time_smoothness_weight, l1_time_planes_weight, plane_tv_weight):
    """``GaussianModel.compute_regulation`` (scene/gaussian_model.py:748-749) over
    ``deformation_net.grid.grids`` (a ModuleList of 6-plane ParameterLists; levels with 3 planes carry no
    regulariser, gaussian_model.py:716-717)."""
    coeffs, planes = [], []
    for grids in multi_res_grids:
        if len(grids) == 3:
            continue
        for k, p in enumerate(grids):
            coeffs.append((plane_tv_weight, 0.0) if k in SPATIAL_PLANES else (time_smoothness_weight, l1_time_planes_weight))
            planes.append(p)
    if not planes:
        return torch.zeros((), device="cuda")
    return _PlaneRegulation.apply(tuple(coeffs), *planes)


def compute_plane_smoothness(t):
    """scene/regulation.py:22-28 for one plane."""
    return _PlaneRegulation.apply(((1.0, 0.0),), t)
