#!/bin/bash
# round-2 GPU call N: preprocess_backward register cap A/B (5 vs 6 vs 8 resident blocks per SM)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for mb in 0 6 8 0 6 8; do
  echo "== S3G_PBWD_MINB=$mb"
  S3G_PBWD_MINB=$mb timeout 600 python bench.py --no-cpu-baseline --no-train-iteration --no-extra-configs 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['stages']['backward_ms'])"
done | tee $O/r02n_pbwd_ab.log
