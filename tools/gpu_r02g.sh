#!/bin/bash
# round-2 GPU call G: source-level ncu capture of the two dominant fine-stage kernels (500 k Gaussians) + scan check
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
DEV_P=500000 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'deform_backward_kernel|hexplane_scatter_kernel|hexplane_sample_kernel|deform_forward_tc' -c 4 -o $O/r02g_fine python tools/dev_deform.py --bwd --time --notest > $O/r02g_fine.log 2>&1
tail -5 $O/r02g_fine.log
echo "== quick raster check after the scan change"
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_deform.py -m gpu -q -x > $O/r02g_tests.log 2>&1; tail -3 $O/r02g_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-train-iteration --no-extra-configs > $O/r02g_bench.json 2>$O/r02g_bench.err; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02g_bench.json') if l.startswith('{')][-1]); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['stages'])"
