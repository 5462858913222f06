"""torchrun check + timing of the peer-memory all-reduce against NCCL (GPU box, N >= 2):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dev_peer.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from s3gaussian_b200 import dp

rank, world, local = dp.init_from_env("nccl")
dev = torch.device(f"cuda:{local}")
numel = int(os.environ.get("NUMEL", 118_000_000))      # 2M Gaussians x 59 floats
par, why = dp.make_peer_all_reduce(numel, dev)
if par is None:
    if rank == 0:
        print(json.dumps({"peer_all_reduce": "unavailable", "why": why}))
    dist.destroy_process_group()
    sys.exit(0)
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for n in (4, 1000 * 4, numel):
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    par.flat(n).copy_(x)
    if n < par.numel:
        par.flat()[n:].zero_()
    out = par.all_reduce_()[:n]
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / (ref.abs().max() + 1e-30))
    # every rank must hold the same bits
    chk = out.double().sum().reshape(1).clone()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(lo.item() == hi.item())
    ok &= err < 1e-6 and same
    if rank == 0:
        print(f"n={n}: rel err vs NCCL {err:.2e}, identical on all ranks: {same}")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


if "--nvls" in sys.argv:      # round-2 path: in-switch reduction through the multicast mapping
    x = torch.randn(numel, device=dev, generator=g)
    ref = x.clone(); dist.all_reduce(ref)
    par.flat(numel).copy_(x)
    out = par.nvls_all_reduce_()[:numel]
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / (ref.abs().max() + 1e-30))
    t_nvls = timeit(par.nvls_all_reduce_)
    if rank == 0:
        print(json.dumps({"nvls_rel_err_vs_nccl": err, "nvls_ms": round(t_nvls, 4)}))

x = torch.randn(numel, device=dev)
t_peer = timeit(par.all_reduce_)
t_nccl = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    gb = numel * 4 / 1e9
    print(json.dumps({"world": world, "bytes": numel * 4, "peer_ms": round(t_peer, 4), "nccl_ms": round(t_nccl, 4),
                      "peer_busbw_GBps": round(2 * (world - 1) / world * gb / (t_peer * 1e-3), 1),
                      "nccl_busbw_GBps": round(2 * (world - 1) / world * gb / (t_nccl * 1e-3), 1), "correct": ok}))
dist.destroy_process_group()
