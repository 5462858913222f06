"""CPU: the train-step oracle (oracle/train_oracle.py) against the golden vectors produced by the REAL
reference code (tools/make_golden_train.py): utils/loss_utils.py, torch.optim.Adam as the reference
constructs it, GaussianModel.add_densification_stats."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import train_oracle as tro          # noqa: E402
import make_golden_train as mg                  # noqa: E402  (input generators only; no reference import)

GOLD = os.path.join(ROOT, "tests", "golden")
LOSS_GOLD = sorted(glob.glob(os.path.join(GOLD, "train_loss_*.npz")))


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max().detach() / (b.abs().max().detach() + 1e-30))


def load_loss_case(path):
    z = np.load(path)
    B, C, H, W = [int(v) for v in z["dims"]]
    img, gt, depth, gt_depth = mg.loss_inputs(B, C, H, W, int(z["seed"]))
    return z, img, gt, depth, gt_depth


@pytest.mark.parametrize("path", LOSS_GOLD, ids=[os.path.basename(p)[11:-4] for p in LOSS_GOLD])
def test_loss_oracle_matches_reference(path):
    z, img, gt, depth, gt_depth = load_loss_case(path)
    img = img.double().requires_grad_(True)
    depth = depth.double().requires_grad_(True)
    l1 = tro.l1_mean(img, gt.double())
    ss = tro.ssim_mean(img, gt.double())
    dl2 = tro.depth_l2(depth, gt_depth.double())
    loss = l1 + 0.5 * dl2 + 0.2 * (1.0 - ss)
    loss.backward()
    assert abs(l1.item() - float(z["l1"])) < 1e-6
    assert abs(ss.item() - float(z["ssim"])) < 1e-5
    assert abs(dl2.item() - float(z["depth_l2"])) < 1e-6
    assert rel(img.grad, z["g_image"]) < 1e-4
    assert rel(depth.grad, z["g_depth"]) < 1e-5


def test_adam_oracle_matches_torch_adam_as_the_reference_builds_it():
    z = np.load(os.path.join(GOLD, "train_adam.npz"))
    params, grads = mg.adam_inputs(int(z["seed"]), int(z["steps"]))
    ps = [p.double() for p in params]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for s, gs in enumerate(grads):
        for i, (name, _, lr) in enumerate(mg.ADAM_SHAPES):
            if name == "nograd":
                continue
            if name == "xyz" and s == 2:
                lr = 1.0e-4
            ps[i], ms[i], vs[i] = tro.adam_step(ps[i], gs[i].double(), ms[i], vs[i], s + 1, lr)
            assert rel(ps[i], z[f"p{i}_s{s}"]) < 2e-6, (name, s)
    for i, (name, _, _) in enumerate(mg.ADAM_SHAPES):
        if name == "nograd":
            assert f"m{i}" not in z.files
            assert np.array_equal(z[f"p{i}_s2"], params[i].numpy())
            continue
        assert rel(ms[i], z[f"m{i}"]) < 2e-6 and rel(vs[i], z[f"v{i}"]) < 2e-6


def test_densify_stats_oracle_matches_reference():
    z = np.load(os.path.join(GOLD, "train_densify_stats.npz"))
    radii, vgrad, accum, denom, max_radii = mg.stats_inputs(int(z["P"]), int(z["seed"]))
    a, d, m = tro.densify_stats(vgrad, radii, accum, denom, max_radii)
    assert rel(a, z["accum"]) < 1e-6
    assert np.array_equal(d.numpy(), z["denom"])
    assert np.array_equal(m.numpy(), z["max_radii2D"])


def test_plane_regulation_oracle_matches_reference():
    z = np.load(os.path.join(GOLD, "train_plane_reg.npz"))
    levels = [[p.double().requires_grad_(True) for p in lv] for lv in mg.reg_inputs(int(z["seed"]))]
    total = tro.compute_regulation(levels, 0.01, 0.0001, 0.0001)
    total.backward()
    assert abs(total.item() - float(z["total"])) < 1e-6 * abs(float(z["total"]))
    for l, lv in enumerate(levels):
        for k, p in enumerate(lv):
            assert rel(p.grad, z[f"g{l}_{k}"]) < 1e-5, (l, k)
