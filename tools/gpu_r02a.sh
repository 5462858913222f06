#!/bin/bash
# round-2 GPU call A: never-run kernels first (under their own timeouts), full GPU suite, both bench arms,
# launch list + ncu --set full of the composite and sort kernels.  Everything lands in gpurun_out/r02a_*.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r02a_gpu.txt 2>&1
echo "== umma mn selftest" ; timeout 180 python tools/dev_umma.py --mn > $O/r02a_umma_mn.log 2>&1 ; echo "rc=$?" >> $O/r02a_umma_mn.log
tail -8 $O/r02a_umma_mn.log
echo "== tcgen05 backward decoder draft vs goldens"
S3G_TC_BWD=1 timeout 600 python -m pytest tests/test_gpu_deform.py -k "golden" -q -x > $O/r02a_tc_bwd.log 2>&1 ; echo "rc=$?" >> $O/r02a_tc_bwd.log
tail -15 $O/r02a_tc_bwd.log
echo "== GPU suite"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 -s > $O/r02a_gputests.log 2>&1 ; echo "rc=$?" >> $O/r02a_gputests.log
grep -E "elementwise|passed|failed|Error|error|assert|FAILED" $O/r02a_gputests.log | tail -60
echo "== bench"
timeout 900 python bench.py > $O/r02a_bench_ours.json 2> $O/r02a_bench_ours.err ; echo "rc=$?"; cat $O/r02a_bench_ours.json; tail -3 $O/r02a_bench_ours.err
timeout 900 python bench.py --impl reference --steps 10 > $O/r02a_bench_ref.json 2> $O/r02a_bench_ref.err ; echo "rc=$?"; cat $O/r02a_bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-iteration > $O/r02a_launches_bench.log 2>&1
python tools/summarize_launches.py $O/r02a_launches.csv 1 > $O/r02a_launches_summary.txt 2>&1 ; head -16 $O/r02a_launches_summary.txt
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'render_|sort_onesweep|sort_histogram|emit_instances|scan_tiles|tile_ranges|preprocess_' -c 16 -o $O/r02a_full python tools/dev_profile.py ours 2000000 1920 1280 sh 1 > $O/r02a_full.log 2>&1
ls -la $O | tail -20
