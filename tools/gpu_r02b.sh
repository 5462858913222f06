#!/bin/bash
# round-2 GPU call B: new binning chain (device-side count, tile histogram), onesweep at 4 blocks/SM, config legs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== GPU suite"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 > $O/r02d_gputests.log 2>&1 ; echo "rc=$?" >> $O/r02d_gputests.log
grep -E "passed|failed|Error|error|assert|FAILED" $O/r02d_gputests.log | tail -40
echo "== bench ours"
timeout 1500 python bench.py > $O/r02d_bench_ours.json 2> $O/r02d_bench_ours.err ; echo "rc=$?"; cat $O/r02d_bench_ours.json; tail -5 $O/r02d_bench_ours.err
echo "== bench reference"
timeout 1500 python bench.py --impl reference --steps 10 --configs config4 > $O/r02d_bench_ref.json 2> $O/r02d_bench_ref.err ; echo "rc=$?"; cat $O/r02d_bench_ref.json; tail -5 $O/r02d_bench_ref.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-iteration --no-extra-configs > $O/r02d_launches_bench.log 2>&1
python tools/summarize_launches.py $O/r02d_launches.csv 1 > $O/r02d_launches_summary.txt 2>&1 ; head -16 $O/r02d_launches_summary.txt
