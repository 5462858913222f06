"""PyTorch cold path of the deformation module for the switches the fused kernels do not implement.

SURVEY.md section 8 keeps three non-default ``ModelHiddenParams`` switches out of the CUDA hot path and in
PyTorch (they are off in every shipped config, arguments/__init__.py:204-233):

    static_mlp      mask = static_mlp(hidden)                       scene/deformation.py:31-32,113-114
    empty_voxel     mask = DenseGrid(1, [64,64,64])(xyz)            scene/deformation.py:29-30,115-116, scene/grid.py:15-41
    apply_rotation  rotations = q (x) dr, normalised                scene/deformation.py:141-144, utils/graphics_utils.py:154-177

With any of them set, ``deform_network`` routes ``forward_dynamic`` / ``render_front`` through the functions below:
the same arithmetic as the reference module written with torch ops on the module's own parameters (HexPlane planes
stay ``[1,32,H,W]`` channels_last; ``F.grid_sample`` reads them as they are), differentiable through autograd.
It is a library path (torch kernels), an order of magnitude slower than the fused kernels, and only exists so
that a checkpoint trained with one of these switches still renders.  ``no_grid`` and ``grid_pe != 0`` stay
unsupported: the reference itself cannot run them (``query_time`` leaves ``hidden`` undefined under ``no_grid``,
scene/deformation.py:80-81,91; ``grid_pe`` sizes ``feature_out`` for 3x the features but feeds it 1x or 5x, :47-50,86-87).
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F


class DenseGrid(nn.Module):
    """scene/grid.py:15-41 (the parts the path touches): a [1,C,X,Y,Z] grid sampled trilinearly."""

    def __init__(self, channels, world_size):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.grid = nn.Parameter(torch.ones([1, channels, *world_size]))

    def set_aabb(self, xyz_max, xyz_min):
        self.register_buffer("xyz_min", torch.Tensor(xyz_min))
        self.register_buffer("xyz_max", torch.Tensor(xyz_max))

    def forward(self, xyz):
        shape = xyz.shape[:-1]
        q = xyz.reshape(1, 1, 1, -1, 3)
        ind_norm = ((q - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        out = F.grid_sample(self.grid, ind_norm, mode="bilinear", align_corners=True)
        return out.reshape(self.channels, -1).T.reshape(*shape, self.channels)


def batch_quaternion_multiply(q1, q2):
    """utils/graphics_utils.py:154-177"""
    w = q1[:, 0] * q2[:, 0] - q1[:, 1] * q2[:, 1] - q1[:, 2] * q2[:, 2] - q1[:, 3] * q2[:, 3]
    x = q1[:, 0] * q2[:, 1] + q1[:, 1] * q2[:, 0] + q1[:, 2] * q2[:, 3] - q1[:, 3] * q2[:, 2]
    y = q1[:, 0] * q2[:, 2] - q1[:, 1] * q2[:, 3] + q1[:, 2] * q2[:, 0] + q1[:, 3] * q2[:, 1]
    z = q1[:, 0] * q2[:, 3] + q1[:, 1] * q2[:, 2] - q1[:, 2] * q2[:, 1] + q1[:, 3] * q2[:, 0]
    q3 = torch.stack((w, x, y, z), dim=1)
    return q3 / torch.norm(q3, dim=1, keepdim=True)


def hexplane_features(grid_module, pts, time):
    """HexPlaneField.get_density (scene/hexplane.py:19-46,73-106,160-175): normalise to the aabb, append t, per level the
    product over the six planes of bilinear samples (align_corners=True, border padding), levels concatenated."""
    aabb = grid_module.aabb
    p = (pts - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0        # normalize_aabb, scene/hexplane.py:19-20
    t = torch.full((p.shape[0], 1), float(time), device=p.device, dtype=p.dtype) if not torch.is_tensor(time) else time
    p4 = torch.cat((p, t.reshape(-1, 1).to(p.dtype)), dim=-1)
    combs = list(itertools.combinations(range(4), 2))
    feats = []
    for planes in grid_module.grids:
        f = 1.0
        for ci, comb in enumerate(combs):
            coords = p4[..., comb].view(1, 1, -1, 2)
            s = F.grid_sample(planes[ci], coords, align_corners=True, mode="bilinear", padding_mode="border")
            f = f * s.view(planes[ci].shape[1], -1).T
        feats.append(f)
    return torch.cat(feats, dim=-1)


def forward_dynamic(net, point, scales, rotations, opacity, shs, time):
    """Deformation.forward_dynamic (scene/deformation.py:108-166) for `net` = our Deformation container.
    -> (pts, scales, rotations, opacity, shs, dx, feat, dshs), raw (pre-activation)."""
    a = net.args
    hidden = net.feature_out(hexplane_features(net.grid, point[:, :3], time))
    if a.static_mlp:
        mask = net.static_mlp(hidden)
    elif a.empty_voxel:
        mask = net.empty_voxel(point[:, :3])
    else:
        mask = torch.ones_like(opacity[:, 0]).unsqueeze(-1)
    if a.no_dx:
        pts, dx = point[:, :3], None
    else:
        dx = net.pos_deform(hidden)
        pts = point[:, :3] * mask + dx
    if a.no_ds:
        sc = scales[:, :3]
    else:
        sc = scales[:, :3] * mask + net.scales_deform(hidden)
    if a.no_dr:
        ro = rotations[:, :4]
    else:
        dr = net.rotations_deform(hidden)
        ro = batch_quaternion_multiply(rotations, dr) if a.apply_rotation else rotations[:, :4] + dr
    if a.no_do:
        op = opacity[:, :1]
    else:
        op = opacity[:, :1] * mask + net.opacity_deform(hidden)
    if a.no_dshs:
        sh, dshs = shs, None
    else:
        dshs = net.shs_deform(hidden).reshape([shs.shape[0], 16, 3])
        sh = shs * mask.unsqueeze(-1) + dshs
    feat = net.dino_head(hidden) if getattr(a, "feat_head", True) else None
    return pts, sc, ro, op, sh, dx, feat, dshs


def sh_to_rgb(shs, xyz, campos, degree):
    """convert_SHs_python (gaussian_renderer/__init__.py:107-117, utils/sh_utils.py:57-112): direction from the
    UNDEFORMED position, clamp_min(+0.5, 0)."""
    K = (degree + 1) ** 2
    d = xyz - campos
    d = d / d.norm(dim=1, keepdim=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    b = [torch.full_like(x, 0.28209479177387814)]
    if K > 1:
        b += [-0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x]
    if K > 4:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.31539156525252005 * (2 * zz - xx - yy),
              -1.0925484305920792 * xz, 0.5462742152960396 * (xx - yy)]
    if K > 9:
        b += [-0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * xy * z,
              -0.4570457994644658 * y * (4 * zz - xx - yy), 0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy),
              -0.4570457994644658 * x * (4 * zz - xx - yy), 1.445305721320277 * z * (xx - yy),
              -0.5900435899266435 * x * (xx - 3 * yy)]
    B = torch.cat(b, dim=1)
    return torch.clamp_min((B.unsqueeze(2) * shs[:, :K, :]).sum(1) + 0.5, 0.0)


def render_front(net, xyz, scaling, rotation, opacity, shs, time, campos, active_sh_degree):
    """the fused front-end's contract (deformation.deform_network.render_front) in torch:
    -> (means3D_final, scales_act, rot_act, opacity_act, colors_precomp, dx, dshs, feat)"""
    pts, sc, ro, op, sh, dx, feat, dshs = forward_dynamic(net, xyz, scaling, rotation, opacity, shs, time)
    colors = sh_to_rgb(sh, xyz, campos.to(xyz.device), int(active_sh_degree))
    return pts, torch.exp(sc), F.normalize(ro), torch.sigmoid(op), colors, dx, dshs, feat
