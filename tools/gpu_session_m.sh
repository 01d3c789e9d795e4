#!/bin/bash
set -u
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-carriers 8 > $OUT/bench_4096.json 2> $OUT/bench_4096.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --no-extra --e2e-carriers 8 > $OUT/bench_512.json 2> $OUT/bench_512.err; echo "bench512 rc=$?" | tee -a $OUT/status.txt
for f in $OUT/bench_4096.json $OUT/bench_512.json; do python tools/bench_line.py $f; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_finalize -c 6 --csv --log-file $OUT/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-carriers 1 > /dev/null 2>&1
grep k_finalize $OUT/fin_launches.csv | awk -F'","' '{print $NF}' | head -6
