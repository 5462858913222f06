"""CPU: the PyTorch cold path of the deformation module (s3gaussian_b200/deformation_cold.py: static_mlp, empty_voxel,
apply_rotation - the switches SURVEY section 8 keeps out of the CUDA kernels) against the REAL reference module
imported from oracle/_ref, on the same state_dict: bit for bit, outputs and gradients.  The product guards these
paths to CUDA tensors like everything else; the arithmetic itself is device-independent torch and is checked here."""
import pytest
import torch

import ref_ext

pytestmark = pytest.mark.skipif(not ref_ext.deform_available(), reason="oracle/_ref/s3g_ref not present")

CASES = [dict(static_mlp=True, no_ds=False, no_dr=False, no_do=False),
         dict(empty_voxel=True, no_ds=False, no_do=False),
         dict(apply_rotation=True, no_dr=False),
         dict(static_mlp=True, apply_rotation=True, no_dr=False, no_dshs=True, feat_head=False)]


@pytest.mark.parametrize("flags", CASES, ids=["static_mlp", "empty_voxel", "apply_rotation", "mixed"])
def test_cold_path_equals_reference_module(flags):
    from s3gaussian_b200 import synthetic as syn
    from s3gaussian_b200 import deformation_cold as dc
    from s3gaussian_b200.deformation import deform_network
    ref_dn, _ = ref_ext.load_ref_deform()
    reso, mres = (16, 12, 10, 7), (1, 2, 4)
    st = syn.make_deform_state(3, reso, mres, weight_scale=0.2)
    ref = ref_dn(ref_ext.ref_deform_args(reso, mres, **flags))
    ref.deformation_net.set_aabb(*[list(a) for a in syn.WAYMO_AABB])
    ref.load_state_dict(st, strict=False)
    ours = deform_network(ref_ext.ref_deform_args(reso, mres, **flags))
    assert ours.deformation_net.cold
    ours.deformation_net.set_aabb(*[list(a) for a in syn.WAYMO_AABB])
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=False)      # incl. static_mlp / empty_voxel.grid
    assert not missing and not unexpected, (missing, unexpected)
    g = torch.Generator().manual_seed(4)
    if flags.get("empty_voxel"):
        v = torch.rand(1, 1, 64, 64, 64, generator=g)
        ref.deformation_net.empty_voxel.grid.data.copy_(v)
        ours.deformation_net.empty_voxel.grid.data.copy_(v)
    P = 150
    lo, hi = torch.tensor(syn.WAYMO_AABB[1]), torch.tensor(syn.WAYMO_AABB[0])
    base = [lo + (hi - lo) * torch.rand(P, 3, generator=g), torch.randn(P, 3, generator=g), torch.randn(P, 4, generator=g),
            torch.randn(P, 1, generator=g), torch.randn(P, 16, 3, generator=g)]
    a_in = [t.clone().requires_grad_(True) for t in base]
    b_in = [t.clone().requires_grad_(True) for t in base]
    a = ref(*a_in, torch.full((P, 1), 0.4))
    b = dc.forward_dynamic(ours.deformation_net, *b_in, 0.4)
    ws = [None if x is None else torch.randn(x.shape, generator=g) for x in a]
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.equal(x, y)
    sum((x * w).sum() for x, w in zip(a, ws) if x is not None).backward()
    sum((y * w).sum() for y, w in zip(b, ws) if y is not None).backward()
    for x, y in zip(a_in, b_in):
        assert (x.grad is None) == (y.grad is None)
        if x.grad is not None:
            assert torch.allclose(x.grad, y.grad, rtol=1e-6, atol=1e-7)
    pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
    checked = 0
    for k, p in pr.items():
        if p.grad is not None:
            assert po[k].grad is not None, k
            assert torch.allclose(p.grad, po[k].grad, rtol=1e-5, atol=1e-7), k
            checked += 1
    assert checked >= 20


def test_unrunnable_upstream_switches_raise_with_the_reason():
    from s3gaussian_b200.deformation import deform_network
    for flag in (dict(no_grid=True), dict(grid_pe=2)):
        with pytest.raises(NotImplementedError, match="reference"):
            deform_network(ref_ext.ref_deform_args((16, 12, 10, 7), (1, 2), **flag))
