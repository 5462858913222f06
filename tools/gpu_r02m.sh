#!/bin/bash
# round-2 GPU call M: no-SSIM fused loss path (tests + both bench arms with the e2e step using each arm's own loss code)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > $O/r02m_tests.log 2>&1; tail -3 $O/r02m_tests.log
for impl in ours reference; do
  timeout 900 python bench.py --impl $impl --no-cpu-baseline --no-train-iteration --no-extra-configs > $O/r02m_bench_$impl.json 2> $O/r02m_bench_$impl.err
  python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02m_bench_$impl.json') if l.startswith('{')][-1]); print('$impl', 'ms', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['what'][-120:])"
done
