"""Timing of the training-step kernels next to what the reference runs (GPU box).
  python tools/dev_train.py [--points 2000000] [--width 1920 --height 1280]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from s3gaussian_b200 import losses, optim

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=2000000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1280)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")


def timeit(fn, iters=a.iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {}
# ---- Adam: the reference's eight groups at P Gaussians + the HexPlane/MLP parameters --------
P = a.points
shapes = [("xyz", (P, 3)), ("f_dc", (P, 1, 3)), ("f_rest", (P, 15, 3)), ("opacity", (P, 1)), ("scaling", (P, 3)), ("rotation", (P, 4))]
reso, mult = [64, 64, 64, 25], [1, 2, 4, 8]
import itertools
for m in mult:
    r = [reso[0] * m, reso[1] * m, reso[2] * m, reso[3]]
    for ca, cb in itertools.combinations(range(4), 2):
        shapes.append(("grid", (1, 32, r[cb], r[ca])))
for _ in range(14):
    shapes.append(("deformation", (64, 64)))
shapes.append(("deformation", (64, 128)))
nel = sum(int(torch.tensor(s).prod()) for _, s in shapes)


def make(opt_cls):
    ps = [torch.nn.Parameter(torch.randn(*s, device=dev) * 0.1) for _, s in shapes]
    groups = {}
    for (n, _), p in zip(shapes, ps):
        groups.setdefault(n, {"params": [], "lr": 1e-3, "name": n})["params"].append(p)
    opt = opt_cls(list(groups.values()), lr=0.0, eps=1e-15)
    for p in ps:
        p.grad = torch.randn_like(p) * 0.01
    return ps, opt


ps, opt = make(optim.FusedAdam)
t_ours = timeit(opt.step)
del ps, opt
torch.cuda.empty_cache()
ps, opt = make(torch.optim.Adam)
t_ref = timeit(opt.step)
del ps, opt
torch.cuda.empty_cache()
out["adam"] = {"elements": nel, "ours_ms": t_ours, "torch_adam_ms": t_ref, "ours_GBps": nel * 28 / t_ours / 1e6,
               "speedup": t_ref / t_ours}

# ---- image loss fwd+bwd ---------------------------------------------------------------------
H, W = a.height, a.width
g = torch.Generator(device=dev).manual_seed(0)
gt = torch.rand(1, 3, H, W, device=dev, generator=g)
img = (gt + 0.1 * torch.randn(1, 3, H, W, device=dev, generator=g)).clamp(0, 1)
gtd = torch.rand(1, 1, H, W, device=dev, generator=g) * 100
dep = (gtd + torch.randn(1, 1, H, W, device=dev, generator=g)).abs()


def ours_loss():
    x = img.clone().requires_grad_(True); d = dep.clone().requires_grad_(True)
    losses.training_loss(x, gt, d, gtd).backward()


t_ours = timeit(ours_loss)
res = {"ours_ms": t_ours}
import ref_ext
if ref_ext.loss_utils_available():
    lu = ref_ext.load_ref_loss_utils()

    def ref_loss():
        x = img.clone().requires_grad_(True); d = dep.clone().requires_grad_(True)
        (lu.l1_loss(x, gt) + 0.5 * lu.compute_depth("l2", d, gtd) + 0.2 * (1.0 - lu.ssim(x, gt))).backward()
    res["reference_ms"] = timeit(ref_loss)
    res["speedup"] = res["reference_ms"] / t_ours
out["image_loss"] = res

# ---- densify stats ------------------------------------------------------------------------------
radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32)
vg = torch.randn(P, 3, device=dev)
acc, den, mr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
t_ours = timeit(lambda: optim.add_densification_stats(vg, radii, acc, den, mr))


def ref_stats():
    vis = radii > 0
    mr[vis] = torch.max(mr[vis], radii[vis])
    acc[vis] += torch.norm(vg[vis, :2], dim=-1, keepdim=True)
    den[vis] += 1


out["densify_stats"] = {"ours_ms": t_ours, "torch_ms": timeit(ref_stats)}
# ---- plane regularisers on the shipped HexPlane ----------------------------------------------
from s3gaussian_b200 import regulation, synthetic as syn
from s3gaussian_b200.deformation import deform_network
from oracle import train_oracle as tro
import ref_ext as _re
net = deform_network(_re.ref_deform_args(syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES)).to(dev)
grids = net.deformation_net.grid.grids


def ours_reg():
    for lv in grids:
        for p in lv:
            p.grad = None
    regulation.compute_regulation(grids, 0.01, 0.0001, 0.0001).backward()


def torch_reg():     # the same torch statements the reference runs (oracle restatement, fp32, on the GPU)
    for lv in grids:
        for p in lv:
            p.grad = None
    tro.compute_regulation(grids, 0.01, 0.0001, 0.0001).backward()


out["plane_regulation"] = {"ours_ms": timeit(ours_reg), "torch_ms": timeit(torch_reg)}

# ---- 3-NN scale initialiser ------------------------------------------------------------------------
from s3gaussian_b200.simple_knn import distCUDA2
import make_golden_knn as mk
pts = mk.knn_inputs("clustered", P, 12).to(dev)
res = {"points": P, "ours_ms": timeit(lambda: distCUDA2(pts), iters=3, warm=1)}
if _re.simple_knn_available():
    ref_knn = _re.load_ref_simple_knn()
    res["reference_ms"] = timeit(lambda: ref_knn(pts), iters=3, warm=1)
    res["bit_exact"] = bool(torch.equal(distCUDA2(pts), ref_knn(pts)))
out["knn_init"] = res

# ---- densify + prune on the model ------------------------------------------------------------------
from s3gaussian_b200.gaussian_model import GaussianModel, default_optimization_params
cl = syn.make_cloud(P, seed=0)
gm = GaussianModel(3).create_from_tensors(cl.xyz.to(dev), cl.features_dc.to(dev), cl.features_rest.to(dev),
                                          cl.scaling.to(dev), cl.rotation.to(dev), cl.opacity.to(dev))
gm.training_setup(default_optimization_params())
for n_ in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation"):
    getattr(gm, n_).grad = torch.zeros_like(getattr(gm, n_))
gm.optimizer.step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gm.xyz_gradient_accum = torch.rand(P, 1, device=dev) * 4e-4
gm.denom = torch.ones(P, 1, device=dev)
torch.cuda.synchronize(); e0.record()
gm.densify(0.0002, 0.005, 30.0, None)
gm.prune(0.0002, 0.005, 30.0, None)
e1.record(); torch.cuda.synchronize()
out["densify_prune"] = {"ours_ms": e0.elapsed_time(e1), "points_before": P, "points_after": int(gm.get_xyz.shape[0])}
print(json.dumps(out))
