#!/bin/bash
# round-2 GPU call F: final evidence - full GPU suite, smoke, both bench arms, launch lists, ncu --set full
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== GPU suite"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 -s > $O/r03b_gputests.log 2>&1 ; echo "rc=$?" >> $O/r03b_gputests.log
grep -E "passed|failed|Error|FAILED" $O/r03b_gputests.log | tail -12
echo "== smoke"
timeout 600 python __graft_entry__.py --smoke > $O/r03b_smoke.log 2>&1 ; echo "rc=$?"; tail -3 $O/r03b_smoke.log
echo "== bench ours"
timeout 1500 python bench.py > $O/r03b_bench_ours.json 2> $O/r03b_bench_ours.err ; echo "rc=$?"; grep "^{" $O/r03b_bench_ours.json | cut -c1-1800
echo "== bench reference"
timeout 1500 python bench.py --impl reference > $O/r03b_bench_ref.json 2> $O/r03b_bench_ref.err ; echo "rc=$?"; grep "^{" $O/r03b_bench_ref.json | cut -c1-900
echo "== ncu launch list (raster)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r03b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-iteration --no-extra-configs > $O/r03b_launches_bench.log 2>&1
python tools/summarize_launches.py $O/r03b_launches.csv 1 > $O/r03b_launches_summary.txt 2>&1 ; head -14 $O/r03b_launches_summary.txt
echo "== ncu launch list (fine stage, config 3 shape)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r03b_launches_fine.csv python bench.py --workload fine --points 500000 --steps 2 --warmup 1 --no-cpu-baseline --no-train-iteration > $O/r03b_launches_fine_bench.log 2>&1
python tools/summarize_launches.py $O/r03b_launches_fine.csv 1 > $O/r03b_launches_fine_summary.txt 2>&1 ; head -16 $O/r03b_launches_fine_summary.txt
echo "== ncu full (raster kernels)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'render_|sort_onesweep|sort_histogram|emit_instances|scan_tiles|tile_offsets|preprocess_' -c 15 -o $O/r03b_full python tools/dev_profile.py ours 2000000 1920 1280 sh 1 > $O/r03b_full.log 2>&1
ls -la $O | grep r03b || true
echo "== deform kernels: golden parity, fwd+bwd time at 500 k (kept activations / recompute) and 2 M, launch list"
DEV_P=500000 timeout 600 python tools/dev_deform.py --bwd --time > $O/r03b_deform_dev.log 2>&1 ; tail -3 $O/r03b_deform_dev.log
DEV_P=500000 timeout 600 python tools/dev_deform.py --bwd --time --notest --recompute 2>&1 | tail -1 | tee -a $O/r03b_deform_dev.log
DEV_P=2000000 timeout 600 python tools/dev_deform.py --bwd --time --notest 2>&1 | tail -1 | tee -a $O/r03b_deform_dev.log
DEV_P=500000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'deform_|hexplane_' -c 72 --csv --log-file $O/r03b_launches_deform.csv python tools/dev_deform.py --bwd --time --notest > $O/r03b_launches_deform.log 2>&1
python tools/summarize_launches.py $O/r03b_launches_deform.csv 1 2>&1 | head -9
DEV_P=500000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'deform_|hexplane_s' -s 6 -c 5 -o $O/r03b_deform_full python tools/dev_deform.py --bwd --time --notest > $O/r03b_deform_full.log 2>&1
ls -la $O | grep r02y
exit 0
