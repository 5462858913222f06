#!/bin/bash
# round-2 GPU call E (2 GPUs): fused gradient exchange - two-rank test, N=2 bench fused vs stand-alone all-reduce
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== two-rank tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r02e_multi.log 2>&1 ; echo "rc=$?" >> $O/r02e_multi.log
grep -E "multi\]|passed|failed|Error|assert" $O/r02e_multi.log | cut -c1-1200 | tail -12
echo "== bench N=2 fused"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02e_bench_n2_fused.json 2> $O/r02e_bench_n2_fused.err ; echo "rc=$?"; tail -3 $O/r02e_bench_n2_fused.err | cut -c1-600
echo "== bench N=2 stand-alone all-reduce"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 --no-fused-exchange > $O/r02e_bench_n2_plain.json 2> $O/r02e_bench_n2_plain.err ; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02e_bench_n2_fused.json", "gpurun_out/r02e_bench_n2_plain.json"):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], j["config"]["collective"][:60])
        for r in j.get("per_rank", []):
            print("   rank", r["rank"], "step", r["step_ms"], "comm", r["comm_ms"], "render", r["render_ms"], "R", r["num_rendered"], "bwd", (r.get("stages") or {}).get("backward_ms"))
        print("   comm", j.get("comm"))
    except Exception as e:
        print(f, "ERR", e)
PY
