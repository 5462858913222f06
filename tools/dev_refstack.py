"""gpurun: how far is render() from the reference stack at 200k (distribution of pixel errors), and how long does
the reference's PyTorch deformation take on this GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, util, ref_ext
from s3gaussian_b200 import synthetic as syn
from s3gaussian_b200.deformation import deform_network
from s3gaussian_b200.gaussian_renderer import render, PipelineParams, GaussianModelLite
DEV = "cuda:0"
ref = ref_ext.load(); ref_deform_network, eval_sh = ref_ext.load_ref_deform()
P, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 960, 640
cloud = syn.make_cloud(P, seed=0, width=W, height=H); cam = syn.make_camera(W, H, (0, 0, 2.0), time=0.37)
st = syn.make_deform_state(0, weight_scale=0.2)
args = ref_ext.ref_deform_args(syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES)
mine = deform_network(args); mine.deformation_net.set_aabb(*[list(a) for a in syn.WAYMO_AABB]); mine.load_state_dict(st, strict=False)
theirs = ref_deform_network(args); theirs.deformation_net.set_aabb(*[list(a) for a in syn.WAYMO_AABB]); theirs.load_state_dict(st, strict=False)
pc = GaussianModelLite(cloud, mine).to(DEV); theirs = theirs.to(DEV)
bg = torch.zeros(3, device=DEV)
out = render(cam.to(DEV), pc, PipelineParams(), bg, stage="fine", return_dx=True, render_feat=True)
xyz = cloud.xyz.to(DEV); sc = cloud.scaling.to(DEV); ro = cloud.rotation.to(DEV); op = cloud.opacity.to(DEV); shs = cloud.get_features().to(DEV)
t = torch.full((P, 1), cam.time, device=DEV)
with torch.no_grad():
    m3, s2, r2, o2, shf, dx, feat, dshs = theirs(xyz, sc, ro, op, shs, t)
    s_a, r_a, o_a = torch.exp(s2), torch.nn.functional.normalize(r2), torch.sigmoid(o2)
    campos = cam.camera_center.to(DEV); dirn = xyz - campos; dirn = dirn / dirn.norm(dim=1, keepdim=True)
    colors = torch.clamp_min(eval_sh(3, shf.transpose(1, 2).view(-1, 3, 16), dirn) + 0.5, 0.0)
    rs = util.settings_for(ref, dict(H=H, W=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, sh_degree=3, campos=cam.camera_center), DEV)
    img, radii, dep = ref.GaussianRasterizer(rs)(means3D=m3, means2D=torch.zeros_like(xyz), shs=None, colors_precomp=colors, opacities=o_a, scales=s_a, rotations=r_a, cov3D_precomp=None)
d = (out["render"].detach() - img).abs()
print("means3D rel", util.relerr(pc._xyz.detach().cpu().numpy()*0 + 0, 0) if False else "", "dx rel", util.relerr(out["dx"].detach().cpu().numpy(), dx.cpu().numpy()))
print("image: max abs", d.max().item(), "mean abs", d.mean().item(), "frac > 1e-4:", (d > 1e-4).float().mean().item(), "frac > 1e-3:", (d > 1e-3).float().mean().item(), "radii mismatches", int((out["radii"] != radii).sum()))
dd = (out["depth"].detach() - dep).abs(); print("depth: max abs", dd.max().item(), "of max", dep.abs().max().item(), "frac > 1e-3:", (dd > 1e-3).float().mean().item())
# reference deformation timing on this GPU (PyTorch ops), fwd and fwd+bwd
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
xg = xyz.clone().requires_grad_(True); sg = shs.clone().requires_grad_(True)
def ref_fwd():
    with torch.no_grad(): theirs(xyz, sc, ro, op, shs, t)
def ref_fb():
    for p in theirs.parameters(): p.grad = None
    o = theirs(xg, sc, ro, op, sg, t); (o[0].sum() + o[5].sum() + o[6].sum() + o[7].sum()).backward()
print(f"reference deform_network on this GPU at P={P}: fwd {timeit(ref_fwd):.2f} ms, fwd+bwd {timeit(ref_fb):.2f} ms")
