#!/bin/bash
# Round-2 GPU visit B: A/B of the grouped fused-kernel launches (finalize beside the next group) at 4096 and 512 carriers.
set -u
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
for G in 0 1 2 4; do
  TETRA_K1_GROUPS=$G timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096_g$G.json 2> $OUT/bench_4096_g$G.err; echo "bench g$G rc=$?" | tee -a $OUT/status.txt
done
for G in 0 1 2; do
  TETRA_K1_GROUPS=$G timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_512_g$G.json 2> $OUT/bench_512_g$G.err; echo "bench512 g$G rc=$?" | tee -a $OUT/status.txt
done
tail -5 $OUT/pytest_gpu.log; cat $OUT/status.txt
for f in $OUT/bench_*.json; do echo $f; python tools/bench_line.py $f; done
