#!/bin/bash
# round-2 GPU call K: onesweep position tickets - per block vs per cluster of 8 vs none (measurement only)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for m in ticket cluster blockidx cluster ticket; do
  echo "== S3G_SORT_MODE=$m"
  S3G_SORT_MODE=$m timeout 600 python bench.py --no-cpu-baseline --no-train-iteration --no-extra-configs 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=j['stages']['forward_ms']; print(j['ms_per_step'], j['e2e']['ms_per_step'], 'depth_sort', f['depth_sort'], 'tile_sort', f['tile_sort'], 'scan', f['scan'])"
done | tee $O/r02k_sort_modes.log
echo "== tests (default mode)"
timeout 1200 python -m pytest tests/test_gpu_raster.py tests/test_gpu_init.py -m gpu -q > $O/r02k_tests.log 2>&1; tail -2 $O/r02k_tests.log
