#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture of the fused kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
tail -c 3000 $OUT/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k1_|k_)" -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --carriers 592 --steps 2 --warmup 1 --no-cpu --e2e-carriers 8 > $OUT/ncu_launch_bench.log 2>&1; echo "ncu-launches rc=$?" | tee -a $OUT/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_ -s 1 -c 1 -o $OUT/k1_full -f \
    python bench.py --carriers 592 --steps 1 --warmup 1 --no-cpu --e2e-carriers 8 > $OUT/ncu_full_bench.log 2>&1; echo "ncu-full rc=$?" | tee -a $OUT/status.txt
fi
tail -5 $OUT/pytest_gpu.log; tail -3 $OUT/smoke.log; cat $OUT/status.txt
