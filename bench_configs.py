"""BASELINE.json configs 1-4 as extra keys of the default bench.py line (N == 1).

Each function times ONE arm - ours (libs3g_b200.so through the drop-in Python API) or the reference (its
unmodified CUDA extension + its own PyTorch modules from oracle/_ref) - on the same synthetic inputs, with CUDA
events after warm-up, and returns a small dict.  `bench.py` calls them after the headline leg; `bench.py --impl
reference` fills the same keys for the reference arm, so the two JSON lines line up key by key.

    config1  reference deform_network forward, 1 k Gaussians, CPU PyTorch (scene/deformation.py; no rasterizer)
             + our fused kernels on the same 1 k inputs (GPU) for scale
    config2  100 k static Gaussians, 960x640, forward-only render, both colour paths (Python-style precomputed
             colours and in-kernel SH)
    config3  500 k Gaussians, fine stage (HexPlane + decoder), 1920x1280, rgb + feat passes, fwd+bwd
    config4  2 M Gaussians, 150-camera ring (50 frames x 3), 1920x1280, training iterations incl. Adam and
             densify + prune every 100 iterations (train.py:489-516), mean and p99 per iteration
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _timed(fn, steps, warmup):
    """median milliseconds per call over `steps` calls, one CUDA event pair per call (a single allocator or
    host hiccup inside a short loop would otherwise dominate the mean)"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]))


# ------------------------------------------------------------------------------------------------ config 1
def config1(impl, dev):
    """P = 1000, default ModelHiddenParams, t = 0.37, CPU, all host threads."""
    from s3gaussian_b200 import synthetic as syn
    import ref_ext
    P = 1000
    out = {"what": "deform_network forward, 1000 Gaussians, default HexPlane (35.8M params)"}
    cloud = syn.make_cloud(P, seed=1)
    state = syn.make_deform_state(0, weight_scale=0.2)
    args = ref_ext.ref_deform_args(syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES)
    t = torch.full((P, 1), 0.37)
    shs = cloud.get_features()
    if ref_ext.deform_available():
        dn, _ = ref_ext.load_ref_deform()
        net = dn(args)
        net.deformation_net.set_aabb(*[list(x) for x in syn.WAYMO_AABB])
        net.load_state_dict(state, strict=False)
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            net(cloud.xyz, cloud.scaling, cloud.rotation, cloud.opacity, shs, t)
            reps = 20
            t0 = time.time()
            for _ in range(reps):
                net(cloud.xyz, cloud.scaling, cloud.rotation, cloud.opacity, shs, t)
            dt = (time.time() - t0) / reps
        out["reference_cpu"] = {"ms": round(dt * 1e3, 3), "gaussians_per_s": round(P / dt, 1),
                                "threads": torch.get_num_threads(), "host_cores": os.cpu_count()}
    if impl == "ours":
        from s3gaussian_b200.deformation import deform_network
        net = deform_network(args)
        net.deformation_net.set_aabb(*[list(x) for x in syn.WAYMO_AABB])
        net.load_state_dict(state, strict=False)
        net = net.to(dev)
        a = [v.to(dev) for v in (cloud.xyz, cloud.scaling, cloud.rotation, cloud.opacity, shs, t)]
        with torch.no_grad():
            ms = _timed(lambda: net(*a), 50, 5)
        out["ours_gpu"] = {"ms": round(ms, 4), "gaussians_per_s": round(P / (ms * 1e-3), 1)}
    return out


# ------------------------------------------------------------------------------------------------ config 2
def config2(impl, mod, dev):
    """100 k static Gaussians, one 960x640 camera, forward only; both colour paths."""
    from s3gaussian_b200 import synthetic as syn
    import util
    P, W, H = 100_000, 960, 640
    cloud = syn.make_cloud(P, seed=0, width=W, height=H)
    cam = syn.make_camera(W, H, (0, 0, 2.0))
    out = {"what": f"{P} static Gaussians, {W}x{H}, forward-only"}
    for mode in ("sh", "rgb"):
        d = util.scene_inputs(cloud, cam, mode=mode, sh_degree=3, bg=(0.0, 0.0, 0.0))
        t = {k: (d[k].to(dev) if d[k] is not None else None) for k in util.TENSOR_KEYS}
        m2d = torch.zeros_like(t["means3D"])
        rast = mod.GaussianRasterizer(util.settings_for(mod, d, dev))

        def fwd():
            with torch.no_grad():
                return rast(means3D=t["means3D"], means2D=m2d, opacities=t["opacities"], shs=t["shs"],
                            colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"],
                            cov3D_precomp=None)
        ms = _timed(fwd, 40, 10)
        radii = fwd()[1]
        out["sh_in_kernel" if mode == "sh" else "precomputed_colours"] = {
            "ms": round(ms, 4), "gaussians_per_s": round(P / (ms * 1e-3), 1), "visible": int((radii > 0).sum())}
    return out


# ------------------------------------------------------------------------------------------------ config 3
def config3(impl, dev, hbm_gbs):
    """500 k Gaussians, render(stage='fine', render_feat=True, return_dx=True), 1920x1280, loss of SURVEY 8d, backward."""
    from s3gaussian_b200 import synthetic as syn
    import ref_ext
    import util
    P, W, H = 500_000, 1920, 1280
    cloud = syn.make_cloud(P, seed=0)
    cam = syn.waymo_ring(W, H, frames=50)[1].to(dev)
    cam.time = 0.37
    state = syn.make_deform_state(0, weight_scale=0.2)
    bg = torch.zeros(3, device=dev)
    g = torch.Generator().manual_seed(3)
    img_d, dep_d, feat_d = (torch.rand(3, H, W, generator=g).to(dev), (torch.rand(1, H, W, generator=g) * 50).to(dev),
                            torch.rand(3, H, W, generator=g).to(dev))
    if impl == "ours":
        from s3gaussian_b200.deformation import deform_network
        from s3gaussian_b200.gaussian_renderer import render, PipelineParams, GaussianModelLite
        net = deform_network(ref_ext.ref_deform_args(syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES))
        net.deformation_net.set_aabb(*[list(x) for x in syn.WAYMO_AABB])
        net.load_state_dict(state, strict=False)
        pc = GaussianModelLite(cloud, net).to(dev)
        leaves = list(pc.parameters())
        pipe = PipelineParams()
        do_render = lambda: render(cam, pc, pipe, bg, stage="fine", return_dx=True, render_feat=True)
    else:
        if not (ref_ext.available() and ref_ext.deform_available()):
            return {"unavailable": "oracle/_ref not built"}
        stack = util.RefFineStack(cloud, state, dev, syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES)
        leaves = stack.leaves()
        do_render = lambda: stack.render(cam, bg, render_feat=True)

    def step():
        for v in leaves:
            v.grad = None
        o = do_render()
        loss = ((o["render"] - img_d).abs().mean() + 0.5 * ((o["depth"] - dep_d) ** 2).mean() +
                0.001 * ((o["feat"] - feat_d) ** 2).mean() + 0.001 * o["dx"].abs().mean() + 0.001 * o["dshs"].abs().mean())
        loss.backward()

    def fwd_only():
        with torch.no_grad():
            do_render()
    ms = _timed(step, 8, 3)
    ms_f = _timed(fwd_only, 8, 2)
    sorted_ms = None
    if impl == "ours":
        # the same step with the Gaussians in Morton order of their positions (GaussianModel.spatial_sort(), a
        # framework feature: consecutive warps of the HexPlane kernels then share texels in L1)
        del pc, leaves
        pc = GaussianModelLite(cloud.morton_sorted(), net).to(dev)
        leaves = list(pc.parameters())
        sorted_ms = (_timed(step, 8, 3), _timed(fwd_only, 8, 2))
    out = {"what": f"{P} Gaussians, fine stage (HexPlane 4 levels x 6 planes x 32 ch + dx/dshs/feat heads), {W}x{H}, rgb + feat "
                   "passes, L1 + depth + feat + dx + dshs loss, backward",
           "ms_fwd_bwd": round(ms, 3), "ms_fwd": round(ms_f, 3), "gaussians_per_s": round(P / (ms * 1e-3), 1)}
    if sorted_ms:
        out["ms_fwd_bwd_spatially_sorted"] = round(sorted_ms[0], 3)
        out["ms_fwd_spatially_sorted"] = round(sorted_ms[1], 3)
    if impl == "ours":
        # B_hex of SURVEY 8d: plane reads + plane-gradient scatter, min(P*12288, 142.9 MB) each, + P*216 B written
        b_hex = 2 * min(P * 12288, 142_909_440) + P * 216
        out["hexplane_algorithmic_bytes"] = int(b_hex)
        out["note"] = "per-kernel times of the fine stage: profiles/ (ncu launch list of this leg)"
    return out


# ------------------------------------------------------------------------------------------------ config 4
def config4(impl, mod, dev, iterations=200):
    """2 M Gaussians over the 150-camera ring: coarse-stage training iterations exactly in train.py's order -
    render, L1 + 0.2 (1 - SSIM) + depth-L2 loss, backward, densification statistics, densify + prune every 100
    iterations (thresholds 2e-4 / 0.005, percent_dense 0.01, train.py:489-516), Adam step, zero_grad."""
    from s3gaussian_b200 import synthetic as syn
    import ref_ext
    P, W, H = 2_000_000, 1920, 1280
    cloud = syn.make_cloud(P, seed=0)
    ring = syn.waymo_ring(W, H, frames=50)
    cams = [c.to(dev) for c in ring]
    order = [1, 0, 2]                                         # dataset_readers.py:619
    cams = [cams[3 * f + o] for f in range(50) for o in order]
    g = torch.Generator().manual_seed(4)
    gt_img = torch.rand(3, H, W, generator=g).to(dev)
    gt_dep = (torch.rand(1, H, W, generator=g) * 50).to(dev)
    bg = torch.zeros(3, device=dev)
    extent = 60.0
    # train.py uses densify_grad_threshold 2e-4 on real images; the synthetic targets give much smaller screen-space
    # gradients, so the threshold is set from a 3-iteration pilot to the 95th percentile of the accumulated
    # gradient norm (~5 % of the visible Gaussians are cloned / split per event); opacity threshold 0.005 as upstream
    thr = {"grad": 0.0002}
    from s3gaussian_b200.gaussian_model import default_optimization_params
    opt = default_optimization_params()
    T = lambda t: t.to(dev)
    if impl == "ours":
        from s3gaussian_b200 import losses
        from s3gaussian_b200.gaussian_model import GaussianModel
        from s3gaussian_b200.gaussian_renderer import render, PipelineParams
        pc = GaussianModel(3).create_from_tensors(T(cloud.xyz), T(cloud.features_dc), T(cloud.features_rest),
                                                  T(cloud.scaling), T(cloud.rotation), T(cloud.opacity))
        pc.active_sh_degree = 3
        pc.training_setup(opt)
        pipe = PipelineParams()

        def iteration(it, cam):
            out = render(cam, pc, pipe, bg, stage="coarse")
            loss = losses.training_loss(out["render"], gt_img, out["depth"], gt_dep)
            loss.backward()
            with torch.no_grad():
                pc.densification_step(out["viewspace_points"].grad, out["radii"])
                if it % 100 == 0:
                    pc.densify(thr["grad"], 0.005, extent, None)
                    pc.prune(thr["grad"], 0.005, extent, None)
            pc.optimizer.step()
            pc.optimizer.zero_grad(set_to_none=True)
        npoints = lambda: int(pc.get_xyz.shape[0])
    else:
        if not (ref_ext.available() and ref_ext.gaussian_model_available()):
            return {"unavailable": "oracle/_ref not built"}
        from s3gaussian_b200.gaussian_model import PARAM_GROUPS, _ATTR
        Ref = ref_ext.load_ref_gaussian_model_class()
        lu = ref_ext.load_ref_loss_utils() if ref_ext.loss_utils_available() else None
        pc = Ref()
        src = {"xyz": cloud.xyz, "f_dc": cloud.features_dc, "f_rest": cloud.features_rest, "opacity": cloud.opacity,
               "scaling": cloud.scaling, "rotation": cloud.rotation}
        for n in PARAM_GROUPS:
            setattr(pc, _ATTR[n], torch.nn.Parameter(T(src[n]).clone().requires_grad_(True)))
        pc.percent_dense = opt.percent_dense
        pc._deformation_table = torch.ones(P, dtype=torch.bool, device=dev)
        lrs = {"xyz": opt.position_lr_init, "f_dc": opt.feature_lr, "f_rest": opt.feature_lr / 20.0,
               "opacity": opt.opacity_lr, "scaling": opt.scaling_lr, "rotation": opt.rotation_lr}
        pc.optimizer = torch.optim.Adam([{"params": [getattr(pc, _ATTR[n])], "lr": lrs[n], "name": n} for n in PARAM_GROUPS],
                                        lr=0.0, eps=1e-15)
        pc.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        pc.denom = torch.zeros(P, 1, device=dev)
        pc.max_radii2D = torch.zeros(P, device=dev)
        pc._deformation_accum = torch.zeros(P, 3, device=dev)

        def iteration(it, cam):
            xyz = pc._xyz
            m2d = torch.zeros_like(xyz, requires_grad=True)
            rs = mod.GaussianRasterizationSettings(
                image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
                viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3,
                campos=cam.camera_center, prefiltered=False, debug=False)
            shs = torch.cat((pc._features_dc, pc._features_rest), dim=1)
            color, radii, depth = mod.GaussianRasterizer(rs)(
                means3D=xyz, means2D=m2d, shs=shs, colors_precomp=None, opacities=torch.sigmoid(pc._opacity),
                scales=torch.exp(pc._scaling), rotations=torch.nn.functional.normalize(pc._rotation), cov3D_precomp=None)
            if lu is not None:
                loss = lu.l1_loss(color, gt_img) + 0.5 * lu.compute_depth("l2", depth, gt_dep) + \
                    0.2 * (1.0 - lu.ssim(color.unsqueeze(0), gt_img.unsqueeze(0)))
            else:
                loss = (color - gt_img).abs().mean() + 0.5 * ((depth - gt_dep) ** 2).mean()
            loss.backward()
            with torch.no_grad():
                vis = radii > 0
                pc.max_radii2D[vis] = torch.max(pc.max_radii2D[vis], radii[vis])
                pc.add_densification_stats(m2d.grad, vis)
                if it % 100 == 0:
                    pc.densify(thr["grad"], 0.005, extent, None, 5, 5, None, None, None)
                    pc.prune(thr["grad"], 0.005, extent, None)
            pc.optimizer.step()
            pc.optimizer.zero_grad(set_to_none=True)
        npoints = lambda: int(pc._xyz.shape[0])

    # steady-state allocator for BOTH arms: the densify events grow every per-Gaussian tensor; a training run
    # that has been going for a while serves those requests from torch's cache, a fresh process pays cudaMalloc
    # (~100 ms per event).  Reserve once, outside the timed loop.
    warm = [torch.empty(1 << 30, dtype=torch.uint8, device=dev) for _ in range(12)]
    del warm
    # ... and load the ~20 torch kernels the densify / prune statements use (CUDA loads modules lazily, a few ms
    # each on first use) on a throw-away 20 k-point model of the same class
    _warm_densify(impl, dev)
    torch.manual_seed(0)
    for w in range(1, 4):
        iteration(w, cams[w % len(cams)])
    torch.cuda.synchronize()
    with torch.no_grad():
        gn = (pc.xyz_gradient_accum / pc.denom.clamp_min(1.0))[pc.denom > 0]
        if gn.numel() > 0:
            thr["grad"] = float(torch.quantile(gn[:: max(1, gn.numel() // 1_000_000)].float(), 0.95))
    n0 = npoints()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iterations + 1)]
    evs[0].record()
    sizes = []
    for it in range(1, iterations + 1):
        iteration(it, cams[it % len(cams)])
        evs[it].record()
        if it % 100 == 0:
            sizes.append(npoints())
    torch.cuda.synchronize()
    ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(iterations)])
    plain = np.array([ms[i] for i in range(iterations) if (i + 1) % 100 != 0])
    events = np.array([ms[i] for i in range(iterations) if (i + 1) % 100 == 0])
    return {"what": f"{P} Gaussians, 150-camera ring (50 frames x 3), {W}x{H}, coarse-stage training iteration: render + L1/SSIM/depth "
                    "loss + backward + stats + Adam, densify + prune at iterations 100 and 200 (thresholds 2e-4 / 0.005)",
            "iterations": iterations, "ms_mean": round(float(ms.mean()), 4), "ms_p50": round(float(np.percentile(ms, 50)), 4),
            "ms_p99": round(float(np.percentile(ms, 99)), 4),
            "ms_mean_without_densify_iterations": round(float(plain.mean()), 4),
            "ms_densify_prune_iterations": [round(float(v), 3) for v in events],
            "densify_grad_threshold": thr["grad"],
            "points_start": n0, "points_after_events": sizes,
            "gaussians_per_s": round(P / (float(ms.mean()) * 1e-3), 1)}


def _warm_densify(impl, dev):
    import ref_ext
    from s3gaussian_b200.gaussian_model import GaussianModel, default_optimization_params, PARAM_GROUPS, _ATTR
    n = 20_000
    g = torch.Generator(device=dev).manual_seed(1)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    opt = default_optimization_params()
    tens = {"xyz": r(n, 3) * 10, "f_dc": r(n, 1, 3), "f_rest": r(n, 15, 3) * 0.1, "opacity": r(n, 1) * 3,
            "scaling": r(n, 3) * 0.9 - 2.0, "rotation": r(n, 4)}
    if impl == "ours":
        m = GaussianModel(3).create_from_tensors(tens["xyz"], tens["f_dc"], tens["f_rest"], tens["scaling"], tens["rotation"],
                                                 tens["opacity"])
        m.training_setup(opt)
    else:
        if not ref_ext.gaussian_model_available():
            return
        m = ref_ext.load_ref_gaussian_model_class()()
        for k in PARAM_GROUPS:
            setattr(m, _ATTR[k], torch.nn.Parameter(tens[k].clone().requires_grad_(True)))
        m.percent_dense = opt.percent_dense
        m._deformation_table = torch.ones(n, dtype=torch.bool, device=dev)
        m.optimizer = torch.optim.Adam([{"params": [getattr(m, _ATTR[k])], "lr": 1e-4, "name": k} for k in PARAM_GROUPS],
                                       lr=0.0, eps=1e-15)
        m.max_radii2D = torch.zeros(n, device=dev)
        m._deformation_accum = torch.zeros(n, 3, device=dev)
    m.xyz_gradient_accum = torch.rand(n, 1, device=dev, generator=g) * 1e-3
    m.denom = torch.ones(n, 1, device=dev)
    for p_ in [getattr(m, _ATTR[k]) for k in PARAM_GROUPS]:
        p_.grad = torch.zeros_like(p_)
    m.optimizer.step()
    with torch.no_grad():
        if impl == "ours":
            m.densify(5e-4, 0.005, 20.0, None)
            m.prune(5e-4, 0.005, 20.0, 20)
        else:
            m.densify(5e-4, 0.005, 20.0, None, 5, 5, None, None, None)
            m.prune(5e-4, 0.005, 20.0, 20)
    torch.cuda.synchronize()


def run_all(impl, mod, dev, hbm_gbs, skip=()):
    out = {}
    for name, fn in (("config1", lambda: config1(impl, dev)), ("config2", lambda: config2(impl, mod, dev)),
                     ("config3", lambda: config3(impl, dev, hbm_gbs)), ("config4", lambda: config4(impl, mod, dev))):
        if name in skip:
            continue
        try:
            t0 = time.time()
            out[name] = fn()
            out[name]["leg_seconds"] = round(time.time() - t0, 1)
        except Exception as ex:      # a broken leg must not cost the headline line
            out[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()
    return out
