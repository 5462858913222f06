"""Generate tests/golden/raster_*.npz from the REAL reference extension.

Runs ON THE GPU BOX (the reference rasterizer is CUDA-only):

    gpurun -- 'python tools/make_golden_raster.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz to tests/golden/ and commit.  Each file holds
the inputs of one rasterizer call and everything the reference produced for it:
colour, depth, radii, the decoded internal state (means2D, conic_opacity, depth,
tiles_touched, sorted point_list + keys, ranges, n_contrib, final_T) and all
gradients for a seeded dL/dout.  The CPU oracle (oracle/splat_oracle.c) and the
CUDA path are both tested against these files.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from s3gaussian_b200 import synthetic as syn
import ref_ext
import util

CASES = [
    # name, P, W, H, mode, sh_degree, seed, cov_precomp
    ("small_sh3", 300, 64, 48, "sh", 3, 1, False),
    ("small_rgb", 300, 64, 48, "rgb", 0, 2, False),
    ("ragged_sh1", 500, 70, 45, "sh", 1, 3, False),       # image not a multiple of 16
    ("covpre_rgb", 200, 48, 32, "rgb", 0, 4, True),       # precomputed 3D covariance path
    ("dense_sh2", 1500, 96, 64, "sh", 2, 5, False),       # saturating pixels (T < 1e-4 early stop)
]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    ref = ref_ext.load()
    for name, P, W, H, mode, deg, seed, covpre in CASES:
        cloud, cam = syn.make_small_scene(P=P, width=W, height=H, seed=seed)
        if name == "dense_sh2":
            cloud.opacity += 2.0
        d = util.scene_inputs(cloud, cam, mode=mode, sh_degree=deg, cov_precomp=covpre)
        gc, gd = util.seeded_grads(d, seed=seed + 100)
        out = util.run_module(ref, d, dev, gc, gd)
        # internal state through the raw extension entry point
        E = torch.Tensor([])
        g = lambda k: d[k].to(dev).contiguous() if d[k] is not None else E
        R, _, _, _, gb, bb, ib = ref._C.rasterize_gaussians(
            d["bg"].to(dev), g("means3D"), g("colors_precomp"), g("opacities"), g("scales"), g("rotations"),
            1.0, g("cov3D_precomp"), d["viewmatrix"].to(dev), d["projmatrix"].to(dev), d["tanfovx"],
            d["tanfovy"], H, W, g("shs"), deg, d["campos"].to(dev), False, False)
        torch.cuda.synchronize()
        rg, rb, ri = ref_ext.decode_geom(gb, P), ref_ext.decode_binning(bb, R), ref_ext.decode_image(ib, W * H)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        save = {f"in_{k}": (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                for k, v in d.items() if v is not None}
        save.update(in_grad_color=gc.numpy(), in_grad_depth=gd.numpy(),
                    color=out["color"].cpu().numpy(), depth=out["depth"].cpu().numpy(),
                    radii=out["radii"].cpu().numpy(), num_rendered=np.int64(R),
                    means2D=rg["means2D"].reshape(P, 2), conic_opacity=rg["conic_opacity"].reshape(P, 4),
                    depths=rg["depths"], tiles_touched=rg["tiles_touched"], cov3D=rg["cov3D"].reshape(P, 6),
                    rgb=rg["rgb"].reshape(P, 3), point_list=rb["point_list"], point_list_keys=rb["point_list_keys"],
                    ranges=ri["ranges"].reshape(-1, 2)[:tiles], n_contrib=ri["n_contrib"], final_T=ri["accum_alpha"])
        for k, v in out["grads"].items():
            save[f"grad_{k}"] = v.cpu().numpy()
        path = os.path.join(outdir, f"raster_{name}.npz")
        np.savez_compressed(path, **save)
        print(name, "R", R, "V", int((out["radii"] > 0).sum()), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
