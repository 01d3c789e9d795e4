#!/bin/bash
# One GPU-box visit, the driver's sequence: parity tests, smoke, bench (N = 1 command), reference arm, freq_offset run, config tool.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --fo-max 5000 > $OUT/bench_fo.json 2> $OUT/bench_fo.err; echo "bench-fo rc=$?" | tee -a $OUT/status.txt
TETRA_U8_CARRIERS=4096 timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/status.txt
tail -4 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/status.txt
python tools/bench_line.py $OUT/bench.json; python tools/bench_line.py $OUT/bench_fo.json; tail -c 600 $OUT/bench_ref.json
