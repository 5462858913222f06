#!/bin/bash
# final confirmation on the committed tree: whole GPU suite, smoke, both bench arms (no profilers)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r03e}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gputests.log 2>&1 ; echo "rc=$?" >> $O/${T}_gputests.log
tail -3 $O/${T}_gputests.log
timeout 600 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1 ; tail -2 $O/${T}_smoke.log
timeout 1200 python bench.py > $O/${T}_bench_ours.json 2> $O/${T}_bench_ours.err ; grep "^{" $O/${T}_bench_ours.json | cut -c1-260
timeout 1200 python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err ; grep "^{" $O/${T}_bench_ref.json | cut -c1-260
exit 0
