"""Two-rank GPU tests of the view-parallel path (NCCL + our peer-memory kernels).  Need >= 2 GPUs on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped loudly on a single-GPU box.

  * peer reduce-scatter / all-gather kernels == ncclAllReduce bit for bit, identical bits on both ranks;
  * the NVLS (multimem.ld_reduce / multimem.st) kernel, where the symmetric allocation has a multicast mapping;
  * a view-parallel render step with the gradient sink (backward writes into the symmetric bucket, no gather
    copy): the all-reduced gradients equal the single-process sum of both views;
  * dp.broadcast_gaussians after a real densify + prune on rank 0; dp.sync_view_stats.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import util
    from s3gaussian_b200 import dp, synthetic as syn
    from s3gaussian_b200 import diff_gaussian_rasterization as ours
    r, w, local = dp.init_from_env("nccl")
    dev = torch.device(f"cuda:{local}")
    out = {}
    # ---- collectives on a ragged buffer -------------------------------------------------------------------
    n = 59 * 40_001 + 4 - (59 * 40_001) % 4
    par, why = dp.make_peer_all_reduce(n, dev)
    out["peer_available"] = par is not None
    out["peer_why"] = why
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    if par is not None:
        par.flat(n).copy_(x)
        got = par.all_reduce_()[:n].clone()
        torch.cuda.synchronize()
        out["peer_equals_nccl"] = bool(torch.equal(got, ref))
        if int(getattr(par.handle, "multicast_ptr", 0) or 0):
            par.flat(n).copy_(x)
            got2 = par.nvls_all_reduce_()[:n].clone()
            torch.cuda.synchronize()
            out["nvls_equals_nccl"] = bool(torch.equal(got2, ref))
            out["nvls_maxerr"] = float((got2 - ref).abs().max())
        else:
            out["nvls_equals_nccl"] = None
    # ---- view-parallel render step through the gradient sink ---------------------------------------------
    P, W, H = 20_000, 320, 208
    cloud = syn.make_cloud(P, seed=1, width=W, height=H)
    cams = [syn.make_camera(W, H, (0, 0, 2.0)), syn.make_camera(W, H, (1.0, 0.5, 2.0), yaw_deg=8)]
    gc, gd = None, None

    def grads_of(cam, sink=None, seed=0):
        d = util.scene_inputs(cloud, cam, mode="sh", sh_degree=3, bg=(0.0, 0.0, 0.0))
        gcv, gdv = util.seeded_grads(d, 50 + seed)
        prev = ours.set_grad_sink(sink)
        try:
            o = util.run_module(ours, d, dev, gcv, gdv)
        finally:
            ours.set_grad_sink(prev)
        return o
    shapes = {"means3D": (P, 3), "shs": (P, 16, 3), "opacities": (P, 1), "scales": (P, 3), "rotations": (P, 4)}
    offs, tot = {}, 0
    for k, s in shapes.items():
        cnt = 1
        for v in s:
            cnt *= v
        offs[k] = (tot, cnt)
        tot += cnt
    par2, _ = dp.make_peer_all_reduce(tot, dev)
    bucket = par2.flat(tot) if par2 is not None else torch.zeros(tot, device=dev)

    def sink(name, shape, device):
        if name not in offs:
            return None
        o, c = offs[name]
        return bucket[o:o + c].view(shape)
    mine = grads_of(cams[rank], sink, seed=rank)
    o0, c0 = offs["means3D"]
    out["sink_aliased"] = bool(mine["grads"]["means3D"].data_ptr() == bucket[o0:o0 + c0].data_ptr())
    if par2 is not None:
        par2.all_reduce_()
    else:
        dist.all_reduce(bucket)
    torch.cuda.synchronize()
    # single-process reference: both views on this rank, summed
    a, b = grads_of(cams[0], None, 0), grads_of(cams[1], None, 1)
    worst = 0.0
    for k, (o, c) in offs.items():
        want = (a["grads"][k].double() + b["grads"][k].double()).reshape(-1)
        got = bucket[o:o + c].double()
        worst = max(worst, float((got - want).abs().max() / (want.abs().max() + 1e-30)))
    out["dp_vs_sequential_relerr"] = worst
    chk = bucket.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["identical_on_all_ranks"] = bool(lo.item() == hi.item())
    # ---- the same step with the exchange fused into the backward kernel (dp.FusedGradExchange) ---------------
    ex, why = dp.make_fused_grad_exchange(shapes, dev)
    out["fused_available"], out["fused_why"] = ex is not None, why
    if ex is not None:
        for rep in range(2):                       # twice: the staging must come back zeroed
            mine2 = grads_of(cams[rank], ex, seed=rank)
            summed = ex.finish()
            torch.cuda.synchronize()
        worst = 0.0
        for k, (o, c, shp) in ex.offsets.items():
            want = (a["grads"][k].double() + b["grads"][k].double()).reshape(-1)
            got = summed[o:o + c].double()
            worst = max(worst, float((got - want).abs().max() / (want.abs().max() + 1e-30)))
        out["fused_vs_sequential_relerr"] = worst
        out["fused_views_alias_bucket"] = bool(mine2["grads"]["shs"].data_ptr() == summed[ex.offsets["shs"][0]:].data_ptr())
        chk = summed.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["fused_identical_on_all_ranks"] = bool(lo.item() == hi.item())
        out["fused_multicast"] = bool(ex.mc)
        # the local (not exchanged) outputs still arrive: screen-space gradient of this rank's view
        out["fused_means2D_ok"] = bool(torch.equal(mine2["grads"]["means2D"], (a if rank == 0 else b)["grads"]["means2D"]) or
                                       float((mine2["grads"]["means2D"] - (a if rank == 0 else b)["grads"]["means2D"]).abs().max()) <
                                       1e-5 * float((a if rank == 0 else b)["grads"]["means2D"].abs().max()))
    # ---- densify on rank 0, broadcast, per-step statistics ---------------------------------------------------
    from s3gaussian_b200.gaussian_model import GaussianModel, default_optimization_params, PARAM_GROUPS, _ATTR
    T = lambda t: t.to(dev)
    pc = GaussianModel(3).create_from_tensors(T(cloud.xyz), T(cloud.features_dc), T(cloud.features_rest), T(cloud.scaling),
                                              T(cloud.rotation), T(cloud.opacity))
    pc.training_setup(default_optimization_params())
    vg = torch.zeros(P, 3, device=dev)
    vg[:, :2] = torch.rand(P, 2, device=dev, generator=g) * 1e-3 * (rank + 1)
    radii = (torch.rand(P, device=dev, generator=g) * 30).to(torch.int32) * (1 if rank == 0 else 2)
    dp.sync_view_stats(vg, radii)
    pc.densification_step(vg, radii)
    out["stats_sum"] = (float(pc.xyz_gradient_accum.double().sum()), float(pc.denom.sum()), float(pc.max_radii2D.max()))
    if rank == 0:
        torch.manual_seed(7)
        pc.densify(5e-4, 0.005, 30.0, None)
        pc.prune(5e-4, 0.005, 30.0, None)
    tensors = {n_: getattr(pc, _ATTR[n_]).data for n_ in PARAM_GROUPS}
    got = dp.broadcast_gaussians(tensors, src=0)
    out["bc_points"] = int(got["xyz"].shape[0])
    sig = torch.stack([got[n_].double().sum() for n_ in sorted(got)])
    lo, hi = sig.clone(), sig.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["bc_identical"] = bool(torch.equal(lo, hi))
    results[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_collectives_sink_and_broadcast(built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    r0, r1 = results[0], results[1]
    print("[multi]", dict(r0))
    print("[multi]", dict(r1))
    for r in (r0, r1):
        assert r["peer_available"], r["peer_why"]
        assert r["peer_equals_nccl"]
        assert r["nvls_equals_nccl"] in (True, None), r.get("nvls_maxerr")
        assert r["dp_vs_sequential_relerr"] < 1e-5
        assert r["fused_available"], r["fused_why"]
        assert r["fused_vs_sequential_relerr"] < 1e-5, r["fused_vs_sequential_relerr"]
        assert r["fused_identical_on_all_ranks"] and r["fused_views_alias_bucket"] and r["fused_means2D_ok"]
        assert r["identical_on_all_ranks"] and r["bc_identical"]
        assert r["bc_points"] > 20_000
    assert r0["stats_sum"] == r1["stats_sum"]
    assert r0["bc_points"] == r1["bc_points"]
