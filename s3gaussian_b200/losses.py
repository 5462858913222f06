"""Image losses of the reference's training step as two stencil kernels.

The reference evaluates, per iteration (train.py:395-419):

    Ll1   = l1_loss(image, gt)                          utils/loss_utils.py:50-51
    ssim_ = ssim(image, gt)                             utils/loss_utils.py:60-96   (5 depthwise 11x11 convs)
    d_l2  = compute_depth("l2", depth, gt_depth)        utils/loss_utils.py:20-45
    loss  = Ll1 + lambda_depth * d_l2 + lambda_dssim * (1 - ssim_) + ...

``image_loss_terms`` returns the three scalars from ONE forward pass (separable 11-tap window in shared
memory; partial sums reduced on the device, no host synchronisation) and one backward pass that
turns autograd's three upstream gradients into dL/dimage and dL/ddepth.  ``l1_loss`` / ``ssim`` /
``compute_depth`` keep the reference's signatures on top of it.  CUDA fp32 only; no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as4(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise RuntimeError(f"{what}: expected [C,H,W] or [B,C,H,W], got {tuple(t.shape)}")
    return t


class _ImageLossTerms(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt_image, depth, gt_depth, max_depth, with_ssim=True):
        lib = _lib.load()
        for t in (image, gt_image, depth, gt_depth):
            if t is not None and (not t.is_cuda or t.dtype != torch.float32):
                raise RuntimeError("image_loss_terms: CUDA float32 tensors only (there is no CPU path)")
        x = y = d = g = None
        if image is not None:
            x = image.contiguous()
            y = gt_image.contiguous()
            if x.shape != y.shape:
                raise RuntimeError(f"image_loss_terms: image {tuple(x.shape)} vs gt {tuple(y.shape)}")
            B, Cc, H, W = x.shape
        else:
            H, W = depth.shape[-2:]
            B, Cc = depth.numel() // (H * W), 0
        if depth is not None:
            d = depth.contiguous()
            g = gt_depth.contiguous()
            if d.numel() != B * H * W or g.numel() != B * H * W:
                raise RuntimeError("image_loss_terms: depth / gt_depth must hold B*H*W values")
        dev = (x if x is not None else d).device
        sums = torch.empty(4, dtype=torch.float64, device=dev)
        ptr = lambda t: t.data_ptr() if t is not None else None
        if with_ssim:
            ws = torch.empty(lib.s3g_image_loss_workspace_bytes(B, Cc, H, W), dtype=torch.uint8, device=dev)
            _lib.check(lib.s3g_image_loss_forward(B, Cc, H, W, ptr(x), ptr(y), ptr(d), ptr(g), float(max_depth),
                                                  sums.data_ptr(), ws.data_ptr(), _stream()), "s3g_image_loss_forward")
        else:       # lambda_dssim == 0: no stencil, no derivative maps
            ws = torch.empty(256 + 16 * 592 * 4, dtype=torch.uint8, device=dev)
            _lib.check(lib.s3g_image_l1_depth_forward(B, Cc, H, W, ptr(x), ptr(y), ptr(d), ptr(g), float(max_depth),
                                                      sums.data_ptr(), ws.data_ptr(), _stream()),
                       "s3g_image_l1_depth_forward")
        ctx.with_ssim = bool(with_ssim)
        n = float(max(B * Cc * H * W, 1))
        l1 = (sums[0] / n).float()
        ss = (sums[1] / n).float()
        dl2 = (sums[2] / sums[3]).float() if d is not None else torch.zeros((), device=dev)
        ctx.save_for_backward(x, y, d, g, sums, ws)
        ctx.dims = (B, Cc, H, W)
        ctx.max_depth = float(max_depth)
        ctx.depth_shape = None if depth is None else depth.shape
        ctx.image_shape = None if image is None else image.shape
        return l1, ss, dl2

    @staticmethod
    def backward(ctx, g_l1, g_ssim, g_dl2):
        x, y, d, g, sums, ws = ctx.saved_tensors
        lib = _lib.load()
        B, Cc, H, W = ctx.dims
        wts = torch.stack([g_l1.reshape(()), g_ssim.reshape(()), g_dl2.reshape(())]).to(torch.float32).contiguous()
        gi = torch.empty_like(x) if x is not None else None
        gd = torch.empty_like(d) if d is not None else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        if ctx.with_ssim:
            _lib.check(lib.s3g_image_loss_backward(B, Cc, H, W, ptr(x), ptr(y), ptr(d), ptr(g), ctx.max_depth,
                                                   wts.data_ptr(), sums.data_ptr(), ws.data_ptr(), ptr(gi), ptr(gd),
                                                   _stream()), "s3g_image_loss_backward")
        else:
            _lib.check(lib.s3g_image_l1_depth_backward(B, Cc, H, W, ptr(x), ptr(y), ptr(d), ptr(g), ctx.max_depth,
                                                       wts.data_ptr(), sums.data_ptr(), ptr(gi), ptr(gd), _stream()),
                       "s3g_image_l1_depth_backward")
        return (gi.view(ctx.image_shape) if gi is not None else None, None,
                gd.view(ctx.depth_shape) if gd is not None else None, None, None, None)


def image_loss_terms(image, gt_image, depth=None, gt_depth=None, max_depth: float = 80.0, with_ssim: bool = True):
    """-> (mean|image-gt|, mean SSIM map, masked depth L2) as 0-d tensors; differentiable w.r.t.
    ``image`` and ``depth``.  image/gt: [C,H,W] or [B,C,H,W] (or both None: depth term only);
    depth/gt_depth: [..., H, W] holding B*H*W values (the reference squeezes them, loss_utils.py:29-30)
    or both None.  with_ssim=False skips the SSIM stencil (the second value is 0), as train.py:417 does when
    lambda_dssim == 0."""
    if (depth is None) != (gt_depth is None) or (image is None) != (gt_image is None):
        raise RuntimeError("image_loss_terms: image/gt_image and depth/gt_depth go in pairs")
    if image is None and depth is None:
        raise RuntimeError("image_loss_terms: nothing to compute")
    x = _as4(image, "image") if image is not None else None
    y = _as4(gt_image, "gt_image") if gt_image is not None else None
    return _ImageLossTerms.apply(x, y, depth, gt_depth, max_depth, with_ssim)


# ---- the reference's function names (utils/loss_utils.py) -----------------------------------
def l1_loss(network_output, gt):
    return image_loss_terms(network_output, gt)[0]


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11 or not size_average:
        raise NotImplementedError("ssim: the fused kernel implements the reference's call (window 11, size_average)")
    return image_loss_terms(img1, img2)[1]


def compute_depth(loss_type, pred_depth, gt_depth, max_depth: float = 80):
    if loss_type != "l2":
        raise NotImplementedError(f"compute_depth: only 'l2' (train.py:411) is fused, got {loss_type!r}")
    return image_loss_terms(None, None, pred_depth, gt_depth, max_depth)[2]


def training_loss(image, gt_image, depth, gt_depth, lambda_dssim=0.2, lambda_depth=0.5, max_depth=80.0):
    """Ll1 + lambda_depth * depth_l2 + lambda_dssim * (1 - ssim): the image part of train.py:395-419."""
    if lambda_dssim == 0:      # train.py:417: `if opt.lambda_dssim != 0` - the SSIM term is not even evaluated
        l1, _, dl2 = image_loss_terms(image, gt_image, depth, gt_depth, max_depth, with_ssim=False)
        return l1 + lambda_depth * dl2
    l1, ss, dl2 = image_loss_terms(image, gt_image, depth, gt_depth, max_depth)
    return l1 + lambda_dssim * (1.0 - ss) + lambda_depth * dl2
