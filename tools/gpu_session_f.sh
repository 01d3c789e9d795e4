#!/bin/bash
set -u
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | grep -v "^\.\+$" | tail -40; cat $OUT/status.txt
