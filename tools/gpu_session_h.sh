#!/bin/bash
set -u
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
for S in 0 1; do
TETRA_EDGE_SERIAL=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 --fo-max 5000 > $OUT/bench_4096_fo_serial$S.json 2> $OUT/bench_fo_$S.err; echo "bench fo serial=$S rc=$?" | tee -a $OUT/status.txt
TETRA_EDGE_SERIAL=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096_serial$S.json 2> $OUT/bench_$S.err; echo "bench serial=$S rc=$?" | tee -a $OUT/status.txt
done
TETRA_CONFIGS=2 timeout 600 python tools/bench_configs.py > $OUT/configs2.json 2> $OUT/configs2.err; echo "configs rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -12; cat $OUT/status.txt
for f in $OUT/bench_4096*.json; do echo $f; python tools/bench_line.py $f; done
cat $OUT/configs2.json
