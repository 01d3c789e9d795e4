#!/bin/bash
# N-GPU visit (gpurun --gpus N -- bash tools/gpu_multi.sh N tag): bench.py under torchrun with the peer-memory exchange and with NCCL
set -u
N=${1:-2}; TAG=${2:-r02_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for G in ${GATHERS:-p2p p2p-fused nccl}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --gather $G > $OUT/bench_$G.json 2> $OUT/bench_$G.err; echo "bench $G rc=$?" | tee -a $OUT/status.txt
tail -3 $OUT/bench_$G.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$G.json").read().strip().splitlines()[-1])
    print("$G", "value %.0f ms/step %.4f k1 %.4f frac %.3f gather_parity %s e2e %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gather_parity"], d["e2e"]["value"]))
except Exception as e: print("no line:", e)
PY
done
