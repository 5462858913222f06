"""Diagnostic: run-to-run spread of one gradient tensor in the reference extension and in ours on the same input.

Both backward passes accumulate per-Gaussian gradients with unordered float atomics, so two runs of the SAME
implementation differ.  This prints, for a (P, W, H, mode, view) case of tests/test_gpu_raster.py, the max-normalised
difference ref-vs-ref, ours-vs-ours and ours-vs-ref over `--runs` runs of each, and the values of the worst element in
every run, so that a parity bar can be set against the reference's own noise floor instead of a guess.

    python tests/diag_grad_spread.py --points 2600000 --mode rgb --view front --key rotations
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2_600_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1280)
    ap.add_argument("--mode", default="rgb")
    ap.add_argument("--view", default="front")
    ap.add_argument("--key", default="rotations")
    ap.add_argument("--runs", type=int, default=4)
    a = ap.parse_args()
    import util
    import ref_ext
    from s3gaussian_b200 import synthetic as syn, diff_gaussian_rasterization as ours
    from test_gpu_raster import bench_camera
    ref = ref_ext.load()
    dev = "cuda:0"
    cloud = syn.make_cloud(a.points, seed=0)
    cam = syn.make_camera(a.width, a.height, (0, 0, 2.0)) if a.view == "front" else bench_camera(a.width, a.height, 0)
    d = util.scene_inputs(cloud, cam, mode=a.mode, sh_degree=3, bg=(0.0, 0.0, 0.0))
    gc, gd = util.seeded_grads(d, 7)
    keys = [a.key] if a.key != "all" else None
    R, M = [], []
    for _ in range(a.runs):
        r = util.run_module(ref, d, dev, gc, gd)["grads"]
        R.append({k: v.cpu().numpy() for k, v in r.items()})
        m = util.run_module(ours, d, dev, gc, gd)["grads"]
        M.append({k: v.cpu().numpy() for k, v in m.items()})
    for k in (keys or list(R[0])):
        mx = float(np.abs(R[0][k]).max())
        rr = max(float(np.abs(R[i][k] - R[j][k]).max()) for i in range(a.runs) for j in range(i))
        mm = max(float(np.abs(M[i][k] - M[j][k]).max()) for i in range(a.runs) for j in range(i))
        rm = [[float(np.abs(M[i][k] - R[j][k]).max()) for j in range(a.runs)] for i in range(a.runs)]
        ref_mean = np.mean([R[i][k].astype(np.float64) for i in range(a.runs)], axis=0)
        our_mean = np.mean([M[i][k].astype(np.float64) for i in range(a.runs)], axis=0)
        print(f"[{k}] tensor max {mx:.3e}  ref-ref {rr / mx:.2e}  ours-ours {mm / mx:.2e}  "
              f"ours-ref max {np.max(rm) / mx:.2e} min {np.min(rm) / mx:.2e}  mean(ours)-mean(ref) "
              f"{float(np.abs(our_mean - ref_mean).max()) / mx:.2e}")
        i, j = np.unravel_index(np.argmax(rm), (a.runs, a.runs))
        idx = np.unravel_index(np.argmax(np.abs(M[i][k] - R[j][k])), R[0][k].shape)
        print("   worst element", tuple(int(x) for x in idx),
              "ref runs", [float(R[q][k][idx]) for q in range(a.runs)],
              "our runs", [float(M[q][k][idx]) for q in range(a.runs)])
        g = int(idx[0])
        print("   gaussian", g, "scale", d["scales"][g].tolist(), "rot", d["rotations"][g].tolist(),
              "opacity", float(d["opacities"][g]), "mean", d["means3D"][g].tolist())


if __name__ == "__main__":
    main()
