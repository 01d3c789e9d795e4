#!/bin/bash
set -u
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 > $OUT/bench_4096.json 2> $OUT/bench_4096.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-carriers 8 --fo-max 5000 > $OUT/bench_4096_fo.json 2> $OUT/bench_4096_fo.err; echo "bench fo rc=$?" | tee -a $OUT/status.txt
TETRA_CONFIGS=u8 TETRA_U8_CARRIERS=1024 timeout 600 python tools/bench_configs.py > $OUT/configs_u8.json 2> $OUT/configs_u8.err; echo "configs rc=$?" | tee -a $OUT/status.txt
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_channelize --launch-skip 1 -c 1 -o $OUT/k1_mode1_full -f \
    python bench.py --carriers 592 --steps 1 --warmup 1 --no-cpu --e2e-carriers 1 --fo-max 5000 > $OUT/ncu_mode1.log 2>&1; echo "ncu mode1 rc=$?" | tee -a $OUT/status.txt
fi
grep -v "^$" $OUT/pytest_gpu.log | tail -12; cat $OUT/status.txt
for f in $OUT/bench_4096.json $OUT/bench_4096_fo.json; do python tools/bench_line.py $f; done
cat $OUT/configs_u8.json
ls -la $OUT
