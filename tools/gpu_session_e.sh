#!/bin/bash
set -u
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096.json 2> $OUT/bench_4096.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_512.json 2> $OUT/bench_512.err; echo "bench512 rc=$?" | tee -a $OUT/status.txt
TETRA_CONFIGS=2,3 timeout 600 python tools/bench_configs.py > $OUT/configs23.json 2> $OUT/configs23.err; echo "configs rc=$?" | tee -a $OUT/status.txt
TETRA_CONFIGS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_config3.csv python tools/bench_configs.py > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?" | tee -a $OUT/status.txt
TETRA_CONFIGS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_config2.csv python tools/bench_configs.py > $OUT/ncu_c2.log 2>&1; echo "ncu c2 rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -30; cat $OUT/status.txt
for f in $OUT/bench_4096.json $OUT/bench_512.json; do python tools/bench_line.py $f; done
cat $OUT/configs23.json
