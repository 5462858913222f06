"""CPU restatement of the per-iteration work either side of render() in the reference's
training loop.  TEST INFRASTRUCTURE ONLY (tier rule 3).

Restates (paths relative to /root/reference):
  l1_loss, gaussian/create_window/_ssim, compute_depth("l2")   utils/loss_utils.py:20-96
  loss composition                                             train.py:395-419
  max_radii2D update + add_densification_stats                 train.py:489-491, scene/gaussian_model.py:693-695
  compute_plane_smoothness, _plane/_time/_l1_regulation,       scene/regulation.py:22-28,
  compute_regulation                                           scene/gaussian_model.py:710-749
Third-party arithmetic restated from its documented behaviour: torch.optim.Adam (call site
scene/gaussian_model.py:189 `torch.optim.Adam(l, lr=0.0, eps=1e-15)`, amsgrad off, no weight decay):
    m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
    p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)

Plain tensor arithmetic (no F.conv2d, no torch.optim) so it runs in float64; pinned against the REAL
reference functions / the real torch.optim.Adam on the CPU (tools/make_golden_train.py ->
tests/golden/train_*.npz).
"""
from __future__ import annotations

import math

import torch


def gaussian_window(window_size=11, sigma=1.5, dtype=torch.float32):
    # loss_utils.py:56-58: float32 tensor of python-double exps, normalised by its own sum
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return (g / g.sum()).to(dtype)


def _blur(img, w1d):
    """zero-padded 11x11 window = outer(w, w) (loss_utils.py:60-64, conv2d padding=5, groups=C)."""
    B, C, H, W = img.shape
    k = w1d.numel()
    r = k // 2
    pad = torch.zeros(B, C, H + 2 * r, W + 2 * r, dtype=img.dtype, device=img.device)
    pad[:, :, r:r + H, r:r + W] = img
    w2d = torch.outer(w1d, w1d).to(img.device)
    out = torch.zeros_like(img)
    for i in range(k):
        for j in range(k):
            out = out + w2d[i, j] * pad[:, :, i:i + H, j:j + W]
    return out


def ssim_mean(img1, img2):
    """loss_utils.py:76-96 with size_average=True."""
    w = gaussian_window(dtype=img1.dtype)
    mu1, mu2 = _blur(img1, w), _blur(img2, w)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _blur(img1 * img1, w) - mu1_sq
    s2 = _blur(img2 * img2, w) - mu2_sq
    s12 = _blur(img1 * img2, w) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def l1_mean(a, b):
    return (a - b).abs().mean()                          # loss_utils.py:50-51


def depth_l2(pred, gt, max_depth=80.0):
    """compute_depth("l2", ...) (loss_utils.py:24-45)."""
    pred, gt = pred.squeeze(), gt.squeeze()
    valid = (gt > 0.01) & (gt < max_depth)
    a = (pred[valid] / max_depth).clamp(0.0, 1.0)
    b = (gt[valid] / max_depth).clamp(0.0, 1.0)
    return ((a - b) ** 2).mean()


def training_loss(image, gt, depth, gt_depth, lambda_dssim=0.2, lambda_depth=0.5, max_depth=80.0):
    """Ll1 + lambda_depth * depth + lambda_dssim * (1 - ssim)   (train.py:395,410-419)."""
    return l1_mean(image, gt) + lambda_depth * depth_l2(depth, gt_depth, max_depth) + \
        lambda_dssim * (1.0 - ssim_mean(image, gt))


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One Adam update, `step` counted from 1.  Returns new (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def densify_stats(viewspace_grad, radii, accum, denom, max_radii2D):
    """train.py:489-491 + gaussian_model.py:693-695; returns new (accum, denom, max_radii2D)."""
    vis = radii > 0
    accum, denom, max_radii2D = accum.clone(), denom.clone(), max_radii2D.clone()
    max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis].to(max_radii2D.dtype))
    accum[vis] = accum[vis] + viewspace_grad[vis, :2].norm(dim=-1, keepdim=True).reshape(accum[vis].shape)
    denom[vis] = denom[vis] + 1
    return accum, denom, max_radii2D


def plane_smoothness(t):
    """scene/regulation.py:22-28: mean squared second difference along dim 2 of [B,C,H,W]."""
    h = t.shape[2]
    first = t[..., 1:, :] - t[..., :h - 1, :]
    second = first[..., 1:, :] - first[..., :h - 2, :]
    return (second ** 2).mean()


def compute_regulation(multi_res_grids, time_smoothness_weight, l1_time_planes_weight, plane_tv_weight):
    """scene/gaussian_model.py:710-749: planes 0,1,3 spatial, 2,4,5 spatio-temporal; 3-plane levels skipped."""
    plane_reg = time_reg = l1_reg = 0.0
    for grids in multi_res_grids:
        if len(grids) == 3:
            continue
        for k in (0, 1, 3):
            plane_reg = plane_reg + plane_smoothness(grids[k])
        for k in (2, 4, 5):
            time_reg = time_reg + plane_smoothness(grids[k])
            l1_reg = l1_reg + (1 - grids[k]).abs().mean()
    return plane_tv_weight * plane_reg + time_smoothness_weight * time_reg + l1_time_planes_weight * l1_reg
