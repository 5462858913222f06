#!/bin/bash
# round-2 GPU call I (4 GPUs): N=4 bench both arms (NCCL path with the gradient sink), plus the >16k-tile test on GPU 0
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02i_gpus.txt
echo "== N=4 ours"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r02i_bench_n4.json 2> $O/r02i_bench_n4.err ; echo "rc=$?"; tail -2 $O/r02i_bench_n4.err | cut -c1-300
echo "== N=4 reference"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 5 --impl reference > $O/r02i_bench_n4_ref.json 2> $O/r02i_bench_n4_ref.err ; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02i_bench_n4.json", "gpurun_out/r02i_bench_n4_ref.json"):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms", j["ms_per_step"], "value", j["value"], "e2e", j["e2e"]["ms_per_step"], (j["config"].get("collective") or "")[:60])
        for r in j.get("per_rank", []):
            print("   rank", r["rank"], "step", r["step_ms"], "comm", r["comm_ms"], "render", r["render_ms"], "V", r["visible"], "R", r["num_rendered"])
        print("   comm", j.get("comm"))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== new raster tests on one GPU"
timeout 900 python -m pytest tests/test_gpu_raster.py -m gpu -q -k "16k or overflow or shared_binning" > $O/r02i_tests.log 2>&1; tail -2 $O/r02i_tests.log
