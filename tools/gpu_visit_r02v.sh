#!/bin/bash
# Visit V: GPU suite, bench (block-end states kernel without the freq_offset registers), sanitizer logs of the final build.
set -u
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
tail -3 $OUT/pytest_gpu.log
for rep in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-carriers 16 > $OUT/bench_$rep.json 2> $OUT/bench_$rep.err
  python tools/bench_line.py $OUT/bench_$rep.json
done
timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --no-extra --e2e-carriers 16 > $OUT/bench512.json 2> /dev/null
python tools/bench_line.py $OUT/bench512.json
bash tools/gpu_sanitize.sh $TAG
cat $OUT/status.txt
