#!/bin/bash
# round-2 GPU call Z (2 GPUs): final tree - two-rank tests, N=2 bench of both arms exactly as the driver launches them
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== two-rank tests"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r02z_multi.log 2>&1 ; echo "rc=$?" >> $O/r02z_multi.log
grep -E "passed|failed|Error|assert" $O/r02z_multi.log | cut -c1-600 | tail -8
echo "== bench N=2 ours"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02z_bench_n2.json 2> $O/r02z_bench_n2.err ; echo "rc=$?"; tail -2 $O/r02z_bench_n2.err | cut -c1-400
echo "== bench N=2 reference"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r02z_bench_n2_ref.json 2> $O/r02z_bench_n2_ref.err ; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02z_bench_n2.json", "gpurun_out/r02z_bench_n2_ref.json"):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], str(j["config"].get("collective"))[:80])
        for r in j.get("per_rank", []):
            print("   rank", r["rank"], "step", r["step_ms"], "comm", r["comm_ms"], "render", r["render_ms"], "R", r["num_rendered"])
    except Exception as e:
        print(f, "ERR", e)
PY
