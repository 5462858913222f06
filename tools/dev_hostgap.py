import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
from s3gaussian_b200 import synthetic as syn, _lib
from s3gaussian_b200 import diff_gaussian_rasterization as ours
dev = torch.device("cuda:0")
P, W, H = 2_000_000, 1920, 1280
cloud = syn.make_cloud(P, seed=0); cam = syn.make_camera(W, H, (0, 0, 2.0))
mode = sys.argv[1] if len(sys.argv) > 1 else "sh"
d = util.scene_inputs(cloud, cam, mode=mode, sh_degree=3, bg=(0, 0, 0))
t = {k: (d[k].to(dev).requires_grad_(True) if d[k] is not None else None) for k in util.TENSOR_KEYS}
m2d = torch.zeros_like(t["means3D"], requires_grad=True)
leaves = [v for v in list(t.values()) + [m2d] if v is not None]
g = torch.Generator(device=dev).manual_seed(1)
gc = torch.randn(3, H, W, device=dev, generator=g); gd = torch.randn(1, H, W, device=dev, generator=g)
rast = ours.GaussianRasterizer(util.settings_for(ours, d, dev))
def step(free=True):
    if free:
        for v in leaves: v.grad = None
    color, radii, depth = rast(means3D=t["means3D"], means2D=m2d, opacities=t["opacities"], shs=t["shs"],
                               colors_precomp=t["colors_precomp"], scales=t["scales"], rotations=t["rotations"], cov3D_precomp=None)
    torch.autograd.backward([color, depth], [gc, gd])
for label, free in (("free grads each step", True),):
    for _ in range(5): step(free)
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    hs = []
    for _ in range(20):
        a = time.perf_counter(); step(free); hs.append(time.perf_counter() - a)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    print(label, mode, "gpu ms/step", e0.elapsed_time(e1) / 20, "wall ms/step", (t1 - t0) * 50, "host-side ms/step", sum(hs) * 50,
          "device allocs", s1["num_device_alloc"] - s0["num_device_alloc"], "device frees", s1["num_device_free"] - s0["num_device_free"],
          "reserved GB", s1["reserved_bytes.all.current"] / 1e9)
# where does host time go? time the pieces of one step with syncs
import cProfile, pstats
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step(True)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
