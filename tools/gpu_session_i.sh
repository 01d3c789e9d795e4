#!/bin/bash
set -u
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_parity.log 2>&1; rc=$?; echo "pytest parity rc=$rc" | tee -a $OUT/status.txt
if [ $rc -ne 0 ]; then grep -v "^$" $OUT/pytest_parity.log | tail -30; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
for S in 0 1; do
TETRA_EDGE_SERIAL=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096_serial$S.json 2> $OUT/bench_$S.err; echo "bench serial=$S rc=$?" | tee -a $OUT/status.txt
TETRA_EDGE_SERIAL=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 --fo-max 5000 > $OUT/bench_4096_fo_serial$S.json 2> $OUT/bench_fo_$S.err; echo "bench fo serial=$S rc=$?" | tee -a $OUT/status.txt
TETRA_EDGE_SERIAL=$S timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_512_serial$S.json 2> $OUT/bench512_$S.err; echo "bench512 serial=$S rc=$?" | tee -a $OUT/status.txt
done
TETRA_K1_GROUPS=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096_groups4.json 2> $OUT/bench_g4.err; echo "bench g4 rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -12; cat $OUT/status.txt
for f in $OUT/bench_*.json; do echo $f; python tools/bench_line.py $f; done
