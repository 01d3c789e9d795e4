#!/bin/bash
# quick GPU visit: parity tests + bench (+ launch list); usage: bash tools/gpu_quick.sh tag [bench args]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.0f MS/s  ms/step %.3f  k1 %.3f ms  frac %.3f  share %.2f  parity %s  e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], d["parity_spot_check"], d["e2e"]["value"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(k1_|k_)" -c 8 --csv --log-file $OUT/launches.csv \
    python bench.py --carriers 592 --steps 2 --warmup 0 --no-cpu --e2e-carriers 8 > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches.csv
