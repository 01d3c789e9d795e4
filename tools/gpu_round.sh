#!/bin/bash
# one GPU session: tests, then bench lines with the measurement switches of the block-end / finalize paths
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err
TETRA_FIN_OVERLAP=0 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_4096_noovl.json 2> /dev/null
TETRA_EDGE_SERIAL=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_4096_serial.json 2> /dev/null
python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_512.json 2> gpurun_out/${TAG}_bench_512.err
TETRA_FIN_OVERLAP=0 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_512_noovl.json 2> /dev/null
TETRA_EDGE_SERIAL=1 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_512_serial.json 2> /dev/null
python bench.py --carriers 4096 --fo-max 12000 --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_4096_fo.json 2> /dev/null
tail -5 gpurun_out/${TAG}_tests.log
for f in gpurun_out/${TAG}_bench_*.json; do echo $f; python tools/bench_line.py $f 2>/dev/null || head -c 600 $f; done
