#!/bin/bash
OUT=gpurun_out/r02_edgemode; mkdir -p $OUT
for C in 4096 512; do
for M in 0 1; do
  TETRA_EDGE_MODE=$M timeout 600 python bench.py --carriers $C --steps 10 --warmup 3 --no-cpu --no-extra --e2e-carriers 8 > $OUT/bench_${C}_mode$M.json 2> $OUT/err_${C}_$M.txt
  echo "C=$C TETRA_EDGE_MODE=$M: $(python tools/bench_line.py $OUT/bench_${C}_mode$M.json)"
done
done
