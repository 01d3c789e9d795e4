#!/bin/bash
set -u
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "u8" > $OUT/pytest_u8.log 2>&1; rc=$?; echo "pytest u8 rc=$rc" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_u8.log | tail -25
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -8
TETRA_CONFIGS=u8 TETRA_U8_CARRIERS=4096 timeout 600 python tools/bench_configs.py > $OUT/configs_u8.json 2> $OUT/configs_u8.err; echo "configs rc=$?" | tee -a $OUT/status.txt
cat $OUT/configs_u8.json; tail -3 $OUT/configs_u8.err
