#!/bin/bash
# Visit T: the whole GPU suite on the final build, smoke, A/B of the finalize kernel's prefetch, the driver's bench line.
set -u
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
tail -4 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
line() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "frac", round(r["frac"], 4), "k1_ms", round(r["kernel_ms"], 4),
      "timeline", r.get("last_step_timeline_ms"), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["u8_ingest"]["value"]))
PY
}
for rep in 1 2; do
  for PF in 1 0; do
    TETRA_FIN_PREFETCH=$PF timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-carriers 16 > $OUT/bench_pf${PF}_$rep.json 2> $OUT/bench_pf${PF}_$rep.err
    line $OUT/bench_pf${PF}_$rep.json
  done
done
for PF in 1 0; do
  TETRA_FIN_PREFETCH=$PF timeout 600 python bench.py --carriers 512 --steps 20 --warmup 3 --no-cpu --no-extra --e2e-carriers 16 > $OUT/bench512_pf${PF}.json 2> /dev/null
  line $OUT/bench512_pf${PF}.json
done
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt
line $OUT/bench.json
cat $OUT/status.txt
