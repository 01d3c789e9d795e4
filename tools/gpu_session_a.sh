#!/bin/bash
# Round-2 GPU visit A: parity tests, smoke, bench (driver command), reference arm, config tool.
set -u
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --fo-max 5000 > $OUT/bench_fo.json 2> $OUT/bench_fo.err; echo "bench-fo rc=$?" | tee -a $OUT/status.txt
TETRA_U8_CARRIERS=1024 timeout 600 python tools/bench_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $OUT/status.txt
tail -15 $OUT/pytest_gpu.log; tail -3 $OUT/smoke.log; cat $OUT/status.txt
tail -c 2500 $OUT/bench.json; tail -c 1500 $OUT/bench_ref.json; tail -c 1500 $OUT/bench_fo.json; cat $OUT/configs.json | head -c 4000
