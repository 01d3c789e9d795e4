#!/bin/bash
set -u
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
for i in 1 2 3; do
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -k config3 > $OUT/pytest_c3_$i.log 2>&1; echo "c3 run $i rc=$?" | tee -a $OUT/status.txt
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -k "config3_wideband_channels" > $OUT/memcheck_c3.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/status.txt
TETRA_CONFIGS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_config3.csv python tools/bench_configs.py > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -30; tail -3 $OUT/pytest_c3_*.log; tail -15 $OUT/memcheck_c3.log; cat $OUT/status.txt
