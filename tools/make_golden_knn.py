"""Generate tests/golden/knn_*.npz with the REAL reference simple-knn extension (oracle/_ref/simple_knn,
built unmodified by oracle/build_ref.sh).  CUDA-only: run on the GPU box,
    gpurun -- python tools/make_golden_knn.py gpurun_out/golden
then copy gpurun_out/golden/knn_*.npz to tests/golden/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

CASES = [("uniform_3k", 3000, 1), ("clustered_5k", 5000, 2), ("duplicates_1k", 1000, 3), ("tiny_5", 5, 4), ("line_2k", 2000, 5)]


def knn_inputs(name, P, seed):
    g = torch.Generator().manual_seed(seed)
    if name.startswith("uniform") or name.startswith("tiny"):
        return torch.rand(P, 3, generator=g) * torch.tensor([100.0, 80.0, 20.0]) - torch.tensor([20.0, 40.0, 5.0])
    if name.startswith("clustered"):      # lidar-like: dense near the origin, sparse far away
        r = torch.rand(P, 1, generator=g) ** 3 * 80.0
        d = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1)
        return d * r * torch.tensor([1.0, 1.0, 0.1])
    if name.startswith("duplicates"):
        base = torch.rand(P // 4, 3, generator=g) * 10
        return base[torch.randint(0, P // 4, (P,), generator=g)]
    if name.startswith("line"):           # degenerate extent along two axes
        x = torch.rand(P, 1, generator=g) * 50
        return torch.cat((x, torch.zeros(P, 1), torch.full((P, 1), 3.0)), dim=1)
    raise KeyError(name)


def main(outdir):
    import ref_ext
    distCUDA2 = ref_ext.load_ref_simple_knn()
    os.makedirs(outdir, exist_ok=True)
    for name, P, seed in CASES:
        pts = knn_inputs(name, P, seed)
        d = distCUDA2(pts.cuda()).cpu().numpy()
        np.savez_compressed(os.path.join(outdir, f"knn_{name}.npz"), name=name, P=P, seed=seed, mean_dist2=d)
        print(name, float(d.mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
