"""A minimal S3Gaussian-style training loop on the B200 path: the control flow of the reference's
train.py (scene_reconstruction, train.py:330-530) with every heavy step running through libs3g_b200.so.

    python examples/train_loop.py [--points 200000] [--iters 300] [--fine]

Synthetic data (there is no dataset reader here): a random cloud initialised through create_from_pcd
(3-NN scales), targets rendered from a perturbed copy of it.  Needs a CUDA device (sm_100a).
"""
import argparse
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from s3gaussian_b200 import losses, synthetic as syn
from s3gaussian_b200.gaussian_model import GaussianModel, default_optimization_params
from s3gaussian_b200.gaussian_renderer import PipelineParams, render


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=200_000)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=640)
    ap.add_argument("--fine", action="store_true", help="train the HexPlane deformation stage too")
    ap.add_argument("--checkpoint", default="", help="write and re-load a reference-format checkpoint here (needs --fine)")
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: s3gaussian_b200 has no CPU path")
    dev = torch.device("cuda:0")
    opt = default_optimization_params()
    cams = [c.to(dev) for c in syn.waymo_ring(a.width, a.height, frames=10)]
    bg = torch.zeros(3, device=dev)
    pipe = PipelineParams()

    # ---- scene: "ground truth" cloud -> target images; training starts from a point cloud of it -------
    truth = syn.make_cloud(a.points, seed=0)
    gt = GaussianModel(3).create_from_tensors(truth.xyz.to(dev), truth.features_dc.to(dev), truth.features_rest.to(dev),
                                              truth.scaling.to(dev), truth.rotation.to(dev), truth.opacity.to(dev))
    gt.active_sh_degree = 3
    with torch.no_grad():
        targets = []
        for c in cams:
            out = render(c, gt, pipe, bg, stage="coarse")
            targets.append((out["render"].clone(), out["depth"].clone()))
    colours = (truth.features_dc[:, 0] * 0.28209479177387814 + 0.5).clamp(0, 1)
    pcd = SimpleNamespace(points=truth.xyz.numpy(), colors=colours.numpy())

    deformation = None
    if a.fine:
        sys.path.insert(0, os.path.join(ROOT, "tests"))      # tests/ref_ext.py holds the ModelHiddenParams defaults
        from ref_ext import ref_deform_args
        from s3gaussian_b200.deformation import deform_network
        deformation = deform_network(ref_deform_args(syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES))
        deformation.deformation_net.set_aabb(*[list(x) for x in syn.WAYMO_AABB])
    gaussians = GaussianModel(3, deformation=deformation).create_from_pcd(pcd, spatial_lr_scale=5.0)   # 3-NN scales
    gaussians.training_setup(opt)
    stage = "fine" if a.fine else "coarse"
    extent = 30.0

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(1, a.iters + 1):
        if it == 11:
            e0.record()
        gaussians.update_learning_rate(it)
        if it % 100 == 0:
            gaussians.oneupSHdegree()
        k = (it - 1) % len(cams)
        out = render(cams[k], gaussians, pipe, bg, stage=stage, return_dx=a.fine, render_feat=False)
        image, depth = out["render"], out["depth"]
        l1, ssim_value, depth_l2 = losses.image_loss_terms(image, targets[k][0], depth, targets[k][1])
        loss = l1 + 0.5 * depth_l2 + 0.2 * (1.0 - ssim_value)                    # train.py:395-419
        if a.fine:
            loss = loss + 0.001 * out["dx"].abs().mean() + 0.001 * out["dshs"].abs().mean() + \
                gaussians.compute_regulation(0.01, 0.0001, 0.0001)
        loss.backward()
        with torch.no_grad():
            gaussians.densification_step(out["viewspace_points"].grad, out["radii"])     # train.py:489-491
            if it > 50 and it % 100 == 0:
                gaussians.densify(0.0002, 0.005, extent, 20 if it > 3000 else None)
                gaussians.prune(0.0002, 0.005, extent, 20 if it > 3000 else None)
                gaussians.spatial_sort()      # ours: keep the model in Morton order (densify appends at the end)
        gaussians.optimizer.step()                                               # one launch over all groups
        gaussians.optimizer.zero_grad(set_to_none=True)
        if it % 50 == 0:
            print(f"iter {it:5d}  loss {loss.item():.5f}  l1 {l1.item():.5f}  ssim {ssim_value.item():.4f}  "
                  f"points {gaussians.get_xyz.shape[0]}")
    e1.record()
    torch.cuda.synchronize()
    # checkpoint in the reference's format (train.py:531: torch.save((gaussians.capture(), iteration), path)) and back
    if a.checkpoint and not a.fine:
        print("--checkpoint needs --fine: capture() stores the deformation network's state_dict")
    elif a.checkpoint:
        torch.save((gaussians.capture(), a.iters), a.checkpoint)
        model_args, first_iter = torch.load(a.checkpoint, weights_only=False)
        resumed = GaussianModel(3, deformation=gaussians._deformation)
        resumed.restore(model_args, opt)
        assert torch.equal(resumed.get_xyz, gaussians.get_xyz) and first_iter == a.iters
        print(f"checkpoint {a.checkpoint}: {resumed.get_xyz.shape[0]} points restored (iteration {first_iter})")
    if a.iters > 10:
        print(f"{e0.elapsed_time(e1) / (a.iters - 10):.3f} ms per iteration over the last {a.iters - 10}")


if __name__ == "__main__":
    main()
