"""CPU, world_size 2 over gloo: the host-side data-parallel logic of s3gaussian_b200/dp.py.
(The render kernels have no CPU path; these tests drive the sharding / bucket / broadcast code
with a small differentiable stand-in for one view's loss.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _view_loss(params, view):
    """stand-in for render(view)+loss: any smooth function of all parameters and the view"""
    xyz, feat, plane = params
    w = torch.tensor([1.0 + view, 0.5 * view, -0.25], dtype=torch.float32)
    return ((xyz * w).sum(1).tanh() * feat.sum(1)).sum() + (plane * (view + 1)).pow(2).mean()


def _make_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn(50, 3, generator=g, requires_grad=True)
    feat = torch.randn(50, 4, generator=g, requires_grad=True)
    plane = torch.randn(1, 8, 5, 6, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    return [xyz, feat, plane]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from s3gaussian_b200 import dp
    r, w, _ = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    views = list(range(4))
    mine = dp.shard_views(views, rank, world)
    assert mine == [v for v in views if v % world == rank]
    params = _make_params()
    bucket = dp.GradBucket(params)
    assert params[2].grad.stride() == params[2].stride()          # channels-last grad view
    bucket.zero()
    for v in mine:
        _view_loss(params, v).backward()                          # accumulates straight into the bucket
    bucket.all_reduce()
    results[f"flat{rank}"] = bucket.flat.clone()
    # densify statistics
    # per-step densify statistics: two steps, each reduced once (train.py:391-392,435-437,489-491)
    acc, den, maxr = torch.zeros(3, 1), torch.zeros(3, 1), torch.zeros(3)
    for step in range(2):
        vgrad = torch.tensor([[3.0, 4.0, 0.0], [0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]) * (rank + 1 + step)
        radii = torch.tensor([2 + rank, 0, 5 * rank + step], dtype=torch.int32)
        dp.sync_view_stats(vgrad, radii)
        vis = radii > 0                                     # what densification_step does on the device
        maxr[vis] = torch.max(maxr[vis], radii[vis].float())
        acc[vis] += torch.norm(vgrad[vis, :2], dim=-1, keepdim=True)
        den[vis] += 1
    results[f"stats{rank}"] = (acc.clone(), den.clone(), maxr.clone())
    # densify on rank 0 changes the point count; everyone must end up with rank 0's tensors
    tensors = {"xyz": torch.arange(21.0).view(7, 3) if rank == 0 else torch.zeros(5, 3),
               "opacity": torch.arange(7.0).view(7, 1) if rank == 0 else torch.zeros(5, 1)}
    out = dp.broadcast_gaussians(tensors, src=0)
    results[f"bc{rank}"] = {k: v.clone() for k, v in out.items()}
    results[f"t{rank}"] = dp.max_over_ranks(3.0 + rank, torch.device("cpu"))
    dist.barrier()
    dist.destroy_process_group()


def test_view_parallel_step_equals_sequential_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    # single-process reference: all 4 views in sequence, gradients summed (train.py:372-392,431)
    params = _make_params()
    for v in range(4):
        _view_loss(params, v).backward()
    ref = torch.cat([params[0].grad.reshape(-1), params[1].grad.reshape(-1),
                     params[2].grad.permute(0, 2, 3, 1).reshape(-1)])     # channels-last memory order
    for r in range(world):
        assert torch.allclose(results[f"flat{r}"], ref, rtol=1e-5, atol=1e-6)
        acc, den, maxr = results[f"stats{r}"]
        # step 0: grads x(1+2)=3, radii max (3, 0, 5); step 1: grads x(2+3)=5, radii (3, 0, 6): norm of the SUMMED
        # gradient once per step and +1 in denom per step - not the sum of per-view norms / visibility counts
        assert torch.allclose(acc, torch.tensor([[5.0 * 3 + 5.0 * 5], [0.0], [3.0 + 5.0]]))
        assert torch.equal(den, torch.tensor([[2.0], [0.0], [2.0]]))
        assert torch.equal(maxr, torch.tensor([3.0, 0.0, 6.0]))
        assert torch.equal(results[f"bc{r}"]["xyz"], torch.arange(21.0).view(7, 3))
        assert results[f"bc{r}"]["opacity"].shape == (7, 1)
        assert results[f"t{r}"] == 4.0


def test_single_process_paths_are_noops():
    from s3gaussian_b200 import dp
    assert dp.shard_views([1, 2, 3], 0, 1) == [1, 2, 3]
    p = [torch.zeros(3, 2, requires_grad=True)]
    b = dp.GradBucket(p)
    (p[0] * 2).sum().backward()
    assert torch.equal(b.all_reduce(), torch.full((6,), 2.0)) and b.nbytes == 24
    t = {"a": torch.ones(2)}
    assert dp.broadcast_gaussians(t) is t
    with pytest.raises(ValueError):
        dp.GradBucket([])
