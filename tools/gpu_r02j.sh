#!/bin/bash
# round-2 GPU call J: does the per-block ticket atomic limit the onesweep passes / the scan?  (blockIdx order = experiment only)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for nt in 0; do
  echo "== S3G_SORT_NO_TICKET=$nt"
  S3G_SORT_NO_TICKET=$nt timeout 600 python bench.py --no-cpu-baseline --no-train-iteration --no-extra-configs 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['stages']['forward_ms'])"
done | tee $O/r02l_ticket_ab.log
echo "== sort + raster tests with tickets (product default)"
timeout 900 python -m pytest tests/test_gpu_raster.py -m gpu -q > $O/r02l_tests.log 2>&1; tail -2 $O/r02l_tests.log
