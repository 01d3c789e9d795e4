#!/bin/bash
# ncu evidence for the bench configuration: launch list (durations), DRAM bytes of every kernel of a step, one full capture of the fused kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
K='regex:k1_channelize|k_edge_|k_finalize|k_pack|k_unpack'
for C in 4096 512; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 16 --csv \
      --log-file gpurun_out/${TAG}_launches_${C}.csv python bench.py --carriers $C --steps 2 --warmup 1 --no-cpu --e2e-carriers 1 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k1_channelize --launch-skip 2 -c 1 -o gpurun_out/${TAG}_k1_full \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-carriers 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edge_states --launch-skip 2 -c 1 -o gpurun_out/${TAG}_kc1_full \
    python bench.py --carriers 512 --steps 2 --warmup 1 --no-cpu --e2e-carriers 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_edge_recursions --launch-skip 2 -c 1 -o gpurun_out/${TAG}_kc2_full \
    python bench.py --carriers 512 --steps 2 --warmup 1 --no-cpu --e2e-carriers 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_finalize --launch-skip 2 -c 1 -o gpurun_out/${TAG}_kf_full \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-carriers 1 > /dev/null 2>&1
ls -la gpurun_out/${TAG}_*
head -40 gpurun_out/${TAG}_launches_512.csv
