#!/bin/bash
# round-2 GPU call H: scatter occupancy A/B (2 vs 3 resident blocks), suite re-check after the flaky-test fix
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for P in 500000 2000000; do
  for mb in 2 3; do
    echo "== deform fwd+bwd P=$P S3G_SCATTER_MINB=$mb"
    S3G_SCATTER_MINB=$mb DEV_P=$P timeout 600 python tools/dev_deform.py --bwd --time --notest 2>&1 | tail -1
  done
done | tee $O/r02h_scatter_ab.log
S3G_SCATTER_MINB=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hexplane_scatter' -c 3 --csv python tools/dev_deform.py --bwd --time --notest 2>/dev/null | grep hexplane | cut -d, -f5,15 | tee -a $O/r02h_scatter_ab.log
S3G_SCATTER_MINB=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hexplane_scatter' -c 3 --csv python tools/dev_deform.py --bwd --time --notest 2>/dev/null | grep hexplane | cut -d, -f5,15 | tee -a $O/r02h_scatter_ab.log
echo "== GPU suite"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 > $O/r02h_gputests.log 2>&1 ; echo "rc=$?" >> $O/r02h_gputests.log
grep -E "passed|failed|FAILED" $O/r02h_gputests.log | tail -5
