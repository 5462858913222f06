"""Run a few fwd+bwd iterations of one implementation (for ncu wrapping).

    python tools/dev_profile.py ours|ref P W H sh|rgb iters
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

from s3gaussian_b200 import synthetic as syn
from s3gaussian_b200 import diff_gaussian_rasterization as ours
import ref_ext
from dev_check import run


def main():
    impl, P, W, H, mode, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6])
    dev = torch.device("cuda:0")
    mod = ours if impl == "ours" else ref_ext.load()
    cloud = syn.make_cloud(P, seed=0).to(dev)
    cam = syn.waymo_ring(W, H, frames=50)[1].to(dev)      # bench.py's rank-0 view (front camera of frame 0)
    bg = torch.zeros(3, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    gc = torch.randn(3, H, W, device=dev, generator=g)
    gd = torch.randn(1, H, W, device=dev, generator=g)
    for _ in range(iters):
        run(mod, cloud, cam, bg, mode, gc, gd)
    torch.cuda.synchronize()
    print("done", impl)


if __name__ == "__main__":
    main()
