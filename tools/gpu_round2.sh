#!/bin/bash
# GPU session: tests + the configs tool with the polyphase front end on and off
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_tests.log
TETRA_U8_CARRIERS=1024 python tools/bench_configs.py > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
TETRA_PFB=0 TETRA_U8_CARRIERS=64 python tools/bench_configs.py > gpurun_out/${TAG}_configs_nopfb.json 2> /dev/null
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err
tail -5 gpurun_out/${TAG}_tests.log
python -c "
import json
for f in ('gpurun_out/${TAG}_configs.json','gpurun_out/${TAG}_configs_nopfb.json'):
    d=json.load(open(f))
    for k,v in d.items(): print(k, v)
"
python tools/bench_line.py gpurun_out/${TAG}_bench_4096.json
