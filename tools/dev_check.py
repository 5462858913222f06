"""Developer check (run under gpurun): ours vs the real reference extension.

    python tools/dev_check.py [P] [W] [H] [sh|rgb]

Prints mismatch counts for every integer output, max relative errors for every
float output / gradient, and CUDA-event timings of both implementations.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from s3gaussian_b200 import _lib, synthetic as syn
from s3gaussian_b200 import diff_gaussian_rasterization as ours
import ref_ext


def settings(mod, cam, bg, sh_degree, debug=False):
    return mod.GaussianRasterizationSettings(
        image_height=cam.image_height, image_width=cam.image_width, tanfovx=cam.tanfovx,
        tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center,
        prefiltered=False, debug=debug)


def run(mod, cloud, cam, bg, mode, gc, gd, keep=None):
    xyz = cloud.xyz.clone().requires_grad_(True)
    scal = cloud.get_scaling().clone().requires_grad_(True)
    rot = cloud.get_rotation().clone().requires_grad_(True)
    opa = cloud.get_opacity().clone().requires_grad_(True)
    m2d = torch.zeros_like(xyz, requires_grad=True)
    shs = cloud.get_features().clone().requires_grad_(True)
    cols = torch.sigmoid(cloud.features_dc[:, 0]).clone().requires_grad_(True)
    rast = mod.GaussianRasterizer(settings(mod, cam, bg, 3))
    kw = dict(means3D=xyz, means2D=m2d, opacities=opa, scales=scal, rotations=rot)
    if mode == "sh":
        kw["shs"] = shs
    else:
        kw["colors_precomp"] = cols
    color, radii, depth = rast(**kw)
    loss = (color * gc).sum() + (depth * gd).sum()
    loss.backward()
    g = dict(xyz=xyz.grad, m2d=m2d.grad, scal=scal.grad, rot=rot.grad, opa=opa.grad,
             feat=(shs.grad if mode == "sh" else cols.grad))
    return color.detach(), radii, depth.detach(), g


def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the reference launches on the legacy default stream; a device sync brackets both
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 960
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 640
    mode = sys.argv[4] if len(sys.argv) > 4 else "sh"
    dev = torch.device("cuda:0")
    print("device", torch.cuda.get_device_name(0), "lib", _lib.load().s3g_build_arch().decode())
    cloud = syn.make_cloud(P, seed=0).to(dev)
    cam = syn.make_camera(W, H, (0, 0, 2.0)).to(dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    gc = torch.randn(3, H, W, device=dev, generator=g)
    gd = torch.randn(1, H, W, device=dev, generator=g)
    ref = ref_ext.load()

    c0, r0, d0, g0 = run(ref, cloud, cam, bg, mode, gc, gd)
    c1, r1, d1, g1 = run(ours, cloud, cam, bg, mode, gc, gd)
    torch.cuda.synchronize()
    V = int((r0 > 0).sum())
    print(f"P={P} {W}x{H} mode={mode} V={V} V/P={V / P:.3f}")
    print("radii mismatches:", int((r0 != r1).sum()))
    print("color rel err:", rel(c1, c0), " depth rel err:", rel(d1, d0))
    print("color max abs:", (c1 - c0).abs().max().item(), "depth max abs:", (d1 - d0).abs().max().item())
    for k in g0:
        print(f"grad {k:5s} rel err: {rel(g1[k], g0[k]):.3e}   (ref max {g0[k].abs().max().item():.3e})")

    # internal state, bit for bit
    xyz = cloud.xyz
    rs = settings(ref, cam, bg, 3)
    args = (bg, xyz, torch.Tensor([]) if mode == "sh" else torch.sigmoid(cloud.features_dc[:, 0]).contiguous(),
            cloud.get_opacity(), cloud.get_scaling(), cloud.get_rotation(), 1.0, torch.Tensor([]),
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, H, W,
            cloud.get_features() if mode == "sh" else torch.Tensor([]), 3, rs.campos, False, False)
    R0, _, _, _, gb, bb, ib = ref._C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    rg, rb, ri = ref_ext.decode_geom(gb, P), ref_ext.decode_binning(bb, R0), ref_ext.decode_image(ib, W * H)

    # ours through the autograd function to get at the buffers
    class Ctx:
        def save_for_backward(self, *a): self.saved = a
        def mark_non_differentiable(self, *a): pass
    ctx = Ctx()
    E = torch.Tensor([])
    o_color, o_radii, o_depth = ours._RasterizeGaussians.forward(
        ctx, xyz, torch.zeros_like(xyz), cloud.get_features() if mode == "sh" else E,
        E if mode == "sh" else torch.sigmoid(cloud.features_dc[:, 0]).contiguous(),
        cloud.get_opacity(), cloud.get_scaling(), cloud.get_rotation(), E, settings(ours, cam, bg, 3))
    torch.cuda.synchronize()
    R1 = ctx.num_rendered
    print("num_rendered ref/ours:", R0, R1)
    gbuf, bbuf, ibuf = ctx.saved[7], ctx.saved[8], ctx.saved[9]

    def view(buf, bid, name, dt):
        off, eb, cnt = _lib.state_field(bid, name, P, R1, W, H)
        base = buf.data_ptr()
        a = ((base + 127) & ~127) - base + off
        raw = buf.cpu().numpy()
        return raw[a:a + eb * cnt].view(dt).copy()

    vis = r0.cpu().numpy() > 0
    xyAB = view(gbuf, 0, "xyAB", np.float32).reshape(P, 4)
    Cod = view(gbuf, 0, "Cod", np.float32).reshape(P, 4)
    rgb = view(gbuf, 0, "rgb", np.float32).reshape(P, 4)
    tt = view(gbuf, 0, "tiles_touched", np.uint32)
    m2 = rg["means2D"].reshape(P, 2); co = rg["conic_opacity"].reshape(P, 4)
    print("tiles_touched mismatches:", int((tt != rg["tiles_touched"]).sum()))
    print("means2D bit mismatches (visible):", int((xyAB[vis, :2].view(np.uint32) != m2[vis].view(np.uint32)).sum()))
    ours_conic = np.stack([xyAB[:, 2], xyAB[:, 3], Cod[:, 0], Cod[:, 1]], 1)
    print("conic_opacity bit mismatches (visible):", int((ours_conic[vis].view(np.uint32) != co[vis].view(np.uint32)).sum()))
    print("depth bit mismatches (visible):", int((rgb[vis, 3].view(np.uint32) != rg["depths"][vis].view(np.uint32)).sum()))
    rr = rg["rgb"].reshape(P, 3)
    print("rgb max abs diff (visible):", float(np.abs(rgb[vis, :3] - rr[vis]).max()), "bit mismatches:",
          int((rgb[vis, :3].view(np.uint32) != rr[vis].view(np.uint32)).sum()))
    if R0 == R1:
        pl = view(bbuf, 1, "point_list", np.uint32)
        plt = view(bbuf, 1, "point_list_tiles", np.uint32)
        print("point_list mismatches:", int((pl != rb["point_list"]).sum()),
              " tile-key mismatches:", int((plt != (rb["point_list_keys"] >> np.uint64(32)).astype(np.uint32)).sum()))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    rng = view(ibuf, 2, "ranges", np.uint32).reshape(-1, 2)
    print("ranges mismatches:", int((rng != ri["ranges"].reshape(-1, 2)[:tiles]).sum()))
    nc = view(ibuf, 2, "n_contrib", np.uint32)
    fT = view(ibuf, 2, "final_T", np.float32)
    print("n_contrib mismatches:", int((nc != ri["n_contrib"]).sum()), "of", W * H,
          " final_T max abs:", float(np.abs(fT - ri["accum_alpha"]).max()),
          " sum n_contrib:", int(ri["n_contrib"].sum()))

    # timings
    def fwd(mod):
        xyz_ = cloud.xyz; rast = mod.GaussianRasterizer(settings(mod, cam, bg, 3))
        kw = dict(means3D=xyz_, means2D=torch.zeros_like(xyz_), opacities=cloud.get_opacity(),
                  scales=cloud.get_scaling(), rotations=cloud.get_rotation())
        if mode == "sh": kw["shs"] = cloud.get_features()
        else: kw["colors_precomp"] = torch.sigmoid(cloud.features_dc[:, 0])
        with torch.no_grad():
            return rast(**kw)
    for name, mod in (("ref", ref), ("ours", ours)):
        tf = timeit(lambda: fwd(mod))
        tfb = timeit(lambda: run(mod, cloud, cam, bg, mode, gc, gd))
        print(f"{name:5s} fwd {tf:8.3f} ms   fwd+bwd(+torch glue) {tfb:8.3f} ms")


if __name__ == "__main__":
    main()
