#!/bin/bash
# Visit S: the chunked host batches and the long-block sync search (tests), e2e with and without chunking, and source-level
# ncu captures of the finalize and block-end-state kernels at the bench configuration.
set -u
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_chunked.py tests/test_gpu_sync.py -m gpu -q -x > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
tail -5 $OUT/pytest_new.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $OUT/bench_chunked.json 2> $OUT/bench_chunked.err; echo "bench rc=$?" | tee -a $OUT/status.txt
TETRA_H2D_CHUNK_MB=-1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $OUT/bench_whole.json 2> $OUT/bench_whole.err; echo "bench-whole rc=$?" | tee -a $OUT/status.txt
for f in $OUT/bench_chunked.json $OUT/bench_whole.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["u8_ingest"]["value"])
PY
done
for MB in 32 64 256; do
  TETRA_H2D_CHUNK_MB=$MB timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $OUT/bench_chunk$MB.json 2> /dev/null
  python - $OUT/bench_chunk$MB.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["e2e"]["value"], d["e2e"]["u8_ingest"]["value"])
PY
done
full() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "$k" --launch-skip $skip -c 1 -o $OUT/$name -f "$@" > $OUT/$name.log 2>&1
  echo "full $name rc=$?" | tee -a $OUT/status.txt
  ncu -i $OUT/$name.ncu-rep --page details > $OUT/${name}_details.txt 2>/dev/null
  ncu -i $OUT/$name.ncu-rep --page source --csv > $OUT/${name}_source.csv 2>/dev/null
}
full kfinalize_4096 regex:k_finalize 2 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-carriers 1
full kstates_4096 regex:k_edge_states 2 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-carriers 1
ls -la $OUT; cat $OUT/status.txt
