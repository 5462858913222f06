"""Generate tests/golden/deform_*.npz by importing the REAL reference module.

Runs in the build container (CPU): scene/deformation.py + scene/hexplane.py of
/root/reference are imported as they are, with stand-ins only for modules the
reference imports but the path never touches (tkinter: `from tkinter import W` at
scene/deformation.py:5; the `scene` / `utils` package __init__s pull plyfile, open3d,
simple_knn - bypassed by registering bare package objects).  Parameters come from
s3gaussian_b200.synthetic.make_deform_state(seed) and are loaded with
load_state_dict, so the files only need to store seeds, inputs and outputs.
"""
import os
import sys
import types
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from s3gaussian_b200 import synthetic as syn


def import_reference():
    sys.modules.setdefault("tkinter", types.ModuleType("tkinter")).W = None
    for name in ("scene", "utils"):
        m = types.ModuleType(name)
        m.__path__ = [f"{REF}/{name}"]
        sys.modules[name] = m
    from scene.deformation import deform_network
    from utils.sh_utils import eval_sh
    return deform_network, eval_sh


def ref_args(resolution, multires, **flags):
    d = dict(net_width=64, timebase_pe=4, defor_depth=1, posebase_pe=10, scale_rotation_pe=2, opacity_pe=2,
             timenet_width=64, timenet_output=32, bounds=1.6, plane_tv_weight=0.0001, time_smoothness_weight=0.01,
             l1_time_planes=0.0001,
             kplanes_config={"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 32,
                             "resolution": list(resolution)},
             multires=list(multires), no_dx=False, no_grid=False, no_ds=True, no_dr=True, no_do=True, no_dshs=False,
             feat_head=True, empty_voxel=False, grid_pe=0, static_mlp=False, apply_rotation=False)
    d.update(flags)
    return Namespace(**d)


CASES = [
    # name, P, resolution, multires, state seed, input seed, time, flags, store_param_grads
    ("default_1k", 1000, syn.DEFAULT_RESOLUTION, syn.DEFAULT_MULTIRES, 0, 1, 0.37, {}, False),      # BASELINE config 1
    ("small_allheads", 300, (16, 12, 10, 7), (1, 2, 4), 3, 4, 0.81, dict(no_ds=False, no_dr=False, no_do=False), True),
    ("small_default", 257, (16, 12, 10, 7), (1, 2), 5, 6, 0.0, {}, True),
    ("small_border", 200, (8, 8, 8, 5), (1, 2), 7, 8, 1.0, {}, True),     # points outside the aabb -> border clamp
]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    deform_network, eval_sh = import_reference()
    for name, P, reso, multires, sseed, iseed, tval, flags, store_pg in CASES:
        st = syn.make_deform_state(sseed, reso, multires, weight_scale=0.2)
        net = deform_network(ref_args(reso, multires, **flags))
        net.deformation_net.set_aabb(*[list(a) for a in syn.WAYMO_AABB])
        missing, unexpected = net.load_state_dict(st, strict=False)
        assert not unexpected and all(k.startswith("timenet") or "_poc" in k for k in missing), (missing, unexpected)
        g = torch.Generator().manual_seed(iseed)
        lo, hi = torch.tensor(syn.WAYMO_AABB[1]), torch.tensor(syn.WAYMO_AABB[0])
        span = 1.3 if name == "small_border" else 1.0
        xyz = (lo + (hi - lo) * (0.5 + span * (torch.rand(P, 3, generator=g) - 0.5))).requires_grad_(True)
        scales = torch.randn(P, 3, generator=g).requires_grad_(True)
        rot = torch.randn(P, 4, generator=g).requires_grad_(True)
        opa = torch.randn(P, 1, generator=g).requires_grad_(True)
        shs = torch.randn(P, 16, 3, generator=g).requires_grad_(True)
        t = torch.full((P, 1), tval)
        m3, sc, ro, op, sh, dx, feat, dshs = net(xyz, scales, rot, opa, shs, t)
        # render() front-end (gaussian_renderer/__init__.py:99-117) with the reference's own eval_sh
        campos = torch.tensor([1.0, -2.0, 2.0])
        sc_a, ro_a, op_a = torch.exp(sc), torch.nn.functional.normalize(ro), torch.sigmoid(op)
        shs_view = sh.transpose(1, 2).view(-1, 3, 16)
        dir_pp = xyz - campos.repeat(P, 1)
        dirn = dir_pp / dir_pp.norm(dim=1, keepdim=True)
        colors = torch.clamp_min(eval_sh(3, shs_view, dirn) + 0.5, 0.0)
        # a scalar loss touching every output with seeded weights -> gradients
        ws = [torch.randn(v.shape, generator=g) for v in (m3, sc_a, ro_a, op_a, colors, dx, feat, dshs)]
        loss = sum((v * w).sum() for v, w in zip((m3, sc_a, ro_a, op_a, colors, dx, feat, dshs), ws))
        loss.backward()
        save = dict(P=P, resolution=np.array(reso), multires=np.array(multires), state_seed=sseed, time=tval,
                    flags=np.array([flags.get("no_ds", True), flags.get("no_dr", True), flags.get("no_do", True)]),
                    campos=campos.numpy(), in_xyz=xyz.detach().numpy(), in_scales=scales.detach().numpy(),
                    in_rot=rot.detach().numpy(), in_opacity=opa.detach().numpy(), in_shs=shs.detach().numpy(),
                    out_means3D=m3.detach().numpy(), out_scales=sc_a.detach().numpy(), out_rot=ro_a.detach().numpy(),
                    out_opacity=op_a.detach().numpy(), out_colors=colors.detach().numpy(), out_dx=dx.detach().numpy(),
                    out_feat=feat.detach().numpy(), out_dshs=dshs.detach().numpy(),
                    g_xyz=xyz.grad.numpy(), g_scales=scales.grad.numpy(), g_rot=rot.grad.numpy(),
                    g_opacity=opa.grad.numpy(), g_shs=shs.grad.numpy())
        for i, w in enumerate(ws):
            save[f"w{i}"] = w.numpy()
        for k, p in net.named_parameters():
            if p.grad is None:
                continue
            gk = "pg_" + k
            if store_pg or "grid" not in k:
                save[gk] = p.grad.numpy()
            else:      # 143 MB of plane gradients: keep moments + a strided sample
                gf = p.grad.reshape(-1).double()
                save[gk + "_sum"] = gf.sum().numpy()
                save[gk + "_abs"] = gf.abs().sum().numpy()
                save[gk + "_sample"] = p.grad.reshape(-1)[::997].numpy()
        path = os.path.join(outdir, f"deform_{name}.npz")
        np.savez_compressed(path, **save)
        print(name, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
