"""CPU: the deformation oracle (oracle/deform_oracle.py) against golden vectors produced by
importing the REAL reference deform_network (tools/make_golden_deform.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import deform_oracle as do
from s3gaussian_b200 import synthetic as syn

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "deform_*.npz")))


def load_deform_case(path):
    z = np.load(path)
    st = syn.make_deform_state(int(z["state_seed"]), tuple(int(v) for v in z["resolution"]),
                               tuple(int(v) for v in z["multires"]), weight_scale=0.2)
    flags = dict(no_ds=bool(z["flags"][0]), no_dr=bool(z["flags"][1]), no_do=bool(z["flags"][2]))
    return z, st, flags


def oracle_outputs(z, st, flags, dtype=torch.float64, need_grad=False):
    T = lambda k: torch.from_numpy(z[k]).to(dtype).requires_grad_(need_grad)
    xyz, scales, rot, opa, shs = T("in_xyz"), T("in_scales"), T("in_rot"), T("in_opacity"), T("in_shs")
    state = {k: v.to(dtype).requires_grad_(need_grad and k != "deformation_net.grid.aabb") for k, v in st.items()}
    t = torch.full((xyz.shape[0], 1), float(z["time"]), dtype=dtype)
    d = do.deform_forward(state, xyz, scales, rot, opa, shs, t, **flags)
    fr = do.render_front(xyz, d, torch.from_numpy(z["campos"]).to(dtype), 3)
    outs = (d["means3D"], fr["scales"], fr["rotations"], fr["opacity"], fr["colors_precomp"], d["dx"], d["feat"], d["dshs"])
    return outs, (xyz, scales, rot, opa, shs), state


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_deform_golden_present():
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[7:-4] for p in GOLD])
def test_deform_oracle_forward(path):
    z, st, flags = load_deform_case(path)
    outs, _, _ = oracle_outputs(z, st, flags)
    names = ("out_means3D", "out_scales", "out_rot", "out_opacity", "out_colors", "out_dx", "out_feat", "out_dshs")
    for o, n in zip(outs, names):
        assert rel(o.detach().numpy(), z[n]) < 2e-5, (n, rel(o.detach().numpy(), z[n]))


@pytest.mark.parametrize("path", [p for p in GOLD if "default_1k" not in p],
                         ids=[os.path.basename(p)[7:-4] for p in GOLD if "default_1k" not in p])
def test_deform_oracle_backward(path):
    z, st, flags = load_deform_case(path)
    outs, leaves, state = oracle_outputs(z, st, flags, need_grad=True)
    loss = sum((o * torch.from_numpy(z[f"w{i}"]).double()).sum() for i, o in enumerate(outs))
    loss.backward()
    for leaf, n in zip(leaves, ("g_xyz", "g_scales", "g_rot", "g_opacity", "g_shs")):
        assert rel(leaf.grad.numpy(), z[n]) < 1e-4, (n, rel(leaf.grad.numpy(), z[n]))
    checked = 0
    for k, v in state.items():
        key = "pg_" + k
        if key in z.files and v.grad is not None:
            assert rel(v.grad.numpy(), z[key]) < 1e-4, (k, rel(v.grad.numpy(), z[key]))
            checked += 1
    assert checked >= 20
