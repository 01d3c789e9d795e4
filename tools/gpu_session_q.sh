#!/bin/bash
set -u
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_helpers.py -m gpu -q -x > $OUT/pytest_a.log 2>&1; rc=$?; echo "pytest parity+helpers rc=$rc" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_a.log | tail -25
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_gpu.log | tail -8
TETRA_CONFIGS=2 timeout 600 python tools/bench_configs.py > $OUT/configs2.json 2> $OUT/configs2.err; echo "configs rc=$?" | tee -a $OUT/status.txt
python - <<PY
import json
d=json.load(open("$OUT/configs2.json"))
print(json.dumps(d.get("config2_other_rates_exact_path"), indent=1))
PY
