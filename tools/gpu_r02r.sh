#!/bin/bash
# round-2 GPU call R: deform kernels after a change - parity tests, golden parity, time, per-kernel launch list
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r02r}
mkdir -p $O
echo "== deform + train GPU tests"
timeout 900 python -m pytest tests/test_gpu_deform.py tests/test_gpu_train.py -m gpu -q -x > $O/${T}_tests.log 2>&1 ; echo "rc=$?"
tail -4 $O/${T}_tests.log
echo "== golden parity + time at 500k"
DEV_P=500000 timeout 600 python tools/dev_deform.py --bwd --time > $O/${T}_dev.log 2>&1 ; tail -8 $O/${T}_dev.log
DEV_P=500000 timeout 600 python tools/dev_deform.py --bwd --time --notest --recompute 2>&1 | tail -1
echo "== per-kernel times (ncu launch list)"
DEV_P=500000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'deform_|hexplane_' -c 72 --csv --log-file $O/${T}_launches.csv python tools/dev_deform.py --bwd --time --notest > $O/${T}_launches.log 2>&1
python tools/summarize_launches.py $O/${T}_launches.csv 1 2>&1 | head -9
if [ "$2" = "ncu" ]; then
echo "== ncu full of the backward decoder"
DEV_P=500000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:deform_backward -s 2 -c 1 -o $O/${T}_dbwd python tools/dev_deform.py --bwd --time --notest > $O/${T}_dbwd.log 2>&1
fi
