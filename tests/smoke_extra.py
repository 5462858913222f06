"""Second half of __graft_entry__.smoke(): one tiny invocation of every other kernel family behind the C ABI on
cuda:0, each checked against its oracle / golden (HexPlane + decoder fwd+bwd incl. the tcgen05 forward, fused
image loss, plane regularisers, multi-tensor Adam, densify stats, row gather, 3-NN)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def run():
    dev = "cuda:0"
    from oracle import knn_oracle, train_oracle as tro
    import make_golden_train as mg
    from test_oracle_deform import GOLD, load_deform_case, rel
    from test_gpu_deform import build_net
    # HexPlane + decoder (4 levels -> tcgen05 forward, mma.sync backward) against the reference golden
    path = [p for p in GOLD if p.endswith("deform_default_1k.npz")][0]
    z, st, flags = load_deform_case(path)
    net = build_net(z, st, flags)
    T = lambda k: torch.from_numpy(z[k]).to(dev).requires_grad_(True)
    xyz, sc, ro, op, shs = T("in_xyz"), T("in_scales"), T("in_rot"), T("in_opacity"), T("in_shs")
    outs = net.render_front(xyz, sc, ro, op, shs, float(z["time"]), torch.from_numpy(z["campos"]).to(dev), 3)
    for o, n in zip(outs, ("out_means3D", "out_scales", "out_rot", "out_opacity", "out_colors", "out_dx", "out_dshs", "out_feat")):
        assert rel(o.detach().cpu().numpy().reshape(z[n].shape), z[n]) < 1e-4, n
    sum((o * o).sum() for o in outs).backward()
    assert torch.isfinite(xyz.grad).all() and float(xyz.grad.abs().max()) > 0
    # fused image loss + plane regularisers
    from s3gaussian_b200 import losses, regulation
    img, gt, depth, gt_depth = [t.to(dev) for t in mg.loss_inputs(1, 3, 37, 53, 11)]
    x = img.clone().requires_grad_(True)
    d = depth.clone().requires_grad_(True)
    ours = losses.training_loss(x, gt, d, gt_depth)
    ours.backward()
    xo, do = img.double().cpu().requires_grad_(True), depth.double().cpu().requires_grad_(True)
    ref = tro.training_loss(xo, gt.double().cpu(), do, gt_depth.double().cpu())
    ref.backward()
    assert abs(ours.item() - ref.item()) < 1e-5 and rel(x.grad.cpu(), xo.grad) < 1e-4
    levels = [[p.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for p in lv] for lv in mg.reg_inputs(41)]
    reg = regulation.compute_regulation(levels, 0.01, 0.0001, 0.0001)
    reg.backward()
    cpu_levels = [[p.detach().cpu().double() for p in lv] for lv in levels]
    assert abs(reg.item() - tro.compute_regulation(cpu_levels, 0.01, 0.0001, 0.0001).item()) < 1e-6
    # Adam, stats, row gather
    from s3gaussian_b200.optim import FusedAdam, add_densification_stats
    p0 = torch.randn(5000, 3, device=dev)
    p = torch.nn.Parameter(p0.clone())
    p.grad = torch.randn_like(p)
    FusedAdam([{"params": [p], "lr": 1e-2, "name": "xyz"}], lr=0.0, eps=1e-15).step()
    want, _, _ = tro.adam_step(p0.double().cpu(), p.grad.double().cpu(), torch.zeros(5000, 3, dtype=torch.float64),
                               torch.zeros(5000, 3, dtype=torch.float64), 1, 1e-2)
    assert rel(p.detach().cpu(), want) < 2e-6
    radii, vgrad, accum, denom, maxr = [t.to(dev) for t in mg.stats_inputs(2000, 31)]
    a_ref, d_ref, m_ref = tro.densify_stats(vgrad.cpu(), radii.cpu(), accum.cpu(), denom.cpu(), maxr.cpu())
    add_densification_stats(vgrad, radii, accum, denom, maxr)
    assert rel(accum.cpu(), a_ref) < 1e-6 and torch.equal(denom.cpu(), d_ref) and torch.equal(maxr.cpu(), m_ref)
    from s3gaussian_b200.gaussian_model import GaussianModel
    g = torch.Generator(device=dev).manual_seed(1)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    m = GaussianModel(3).create_from_tensors(r(300, 3), r(300, 1, 3), r(300, 15, 3), r(300, 3), r(300, 4), r(300, 1))
    keep = torch.arange(300, device=dev) % 3 != 0
    before = m._features_rest.data[keep].clone()
    m.prune_points(~keep)
    assert torch.equal(m._features_rest.data, before)
    # spatial sort (row gather + our radix sort): a permutation of every per-Gaussian tensor
    xyz_before = m._xyz.data.clone()
    order = m.spatial_sort()
    assert torch.equal(m._xyz.data, xyz_before[order]) and torch.equal(torch.sort(order).values, torch.arange(order.numel(), device=dev))
    # colour + feature image on one binning (s3g_rasterize_forward_aux / _backward_aux) == two passes
    import util
    from s3gaussian_b200 import synthetic as syn, diff_gaussian_rasterization as dgr
    cloud, cam = syn.make_small_scene(P=500, width=96, height=64, seed=6)
    dd = util.scene_inputs(cloud, cam, mode="rgb")
    rast = dgr.GaussianRasterizer(util.settings_for(dgr, dd, dev))
    tt = {k: dd[k].to(dev).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "colors_precomp")}
    feat = torch.rand(500, 3, device=dev).requires_grad_(True)
    m2d = torch.zeros(500, 3, device=dev, requires_grad=True)
    kw = dict(means3D=tt["means3D"], means2D=m2d, opacities=tt["opacities"], scales=tt["scales"], rotations=tt["rotations"])
    c1, r1, d1, a1 = rast.forward_aux(colors_aux=feat, colors_precomp=tt["colors_precomp"], **kw)
    (c1.sum() + a1.sum() * 0.5 + d1.sum() * 0.1).backward()
    g_fused = feat.grad.clone()
    a2, _, _ = rast(colors_precomp=feat.detach(), **kw)
    assert rel(a1.detach().cpu(), a2.detach().cpu()) < 1e-6 and float(g_fused.abs().max()) > 0
    # 3-NN scale initialiser
    from s3gaussian_b200.simple_knn import distCUDA2
    pts = torch.rand(700, 3, generator=torch.Generator().manual_seed(5)) * 10
    assert np.allclose(distCUDA2(pts.to(dev)).cpu().numpy(), knn_oracle.mean_dist2(pts.numpy()), rtol=2e-6)
    print("smoke extra ok: deform fwd/bwd (tcgen05 forward), image loss, plane regularisers, Adam, stats, row gather, spatial sort, shared-binning aux pass, 3-NN")
