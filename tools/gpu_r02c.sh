#!/bin/bash
# round-2 GPU call C (2 GPUs): two-rank tests, NVLS kernel, N=2 bench both arms, tcgen05 descriptor probe
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02c_gpus.txt
echo "== two-rank tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r02c_multi.log 2>&1 ; echo "rc=$?" >> $O/r02c_multi.log
grep -E "multi\]|passed|failed|Error|assert" $O/r02c_multi.log | cut -c1-600 | tail -12
echo "== peer vs nccl vs nvls"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dev_peer.py --nvls > $O/r02c_peer.log 2>&1 ; echo "rc=$?" >> $O/r02c_peer.log
tail -12 $O/r02c_peer.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02c_bench_n2.json 2> $O/r02c_bench_n2.err ; echo "rc=$?"; cat $O/r02c_bench_n2.json | cut -c1-3000; tail -3 $O/r02c_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 5 --impl reference > $O/r02c_bench_n2_ref.json 2> $O/r02c_bench_n2_ref.err ; echo "rc=$?"; cat $O/r02c_bench_n2_ref.json | cut -c1-1500
echo "== umma probe"
timeout 300 python tools/dev_umma.py --probe > $O/r02c_umma_probe.log 2>&1 ; echo "rc=$?" >> $O/r02c_umma_probe.log
grep -c "==" $O/r02c_umma_probe.log
