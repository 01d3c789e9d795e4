#!/bin/bash
# Final-build ncu evidence: launch lists with DRAM bytes at 4096 / 512 carriers, full captures of the block-end states kernel
# and the finalize kernel (the two kernels that changed after gpu_profile_r02.sh's captures).
set -u
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
K='regex:k1_channelize|k_edge_|k_finalize|k_gather|k_pfb96'
for C in 4096 512; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 24 --csv \
      --log-file $OUT/launches_${C}.csv python bench.py --carriers $C --steps 3 --warmup 1 --no-cpu --no-extra --e2e-carriers 1 > /dev/null 2>&1
  echo "launches $C rc=$?" | tee -a $OUT/status.txt
done
full() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k "$k" --launch-skip $skip -c 1 -o $OUT/$name -f "$@" > $OUT/$name.log 2>&1
  echo "full $name rc=$?" | tee -a $OUT/status.txt
  ncu -i $OUT/$name.ncu-rep --page details > $OUT/${name}_details.txt 2>/dev/null
}
full kstates_4096 regex:k_edge_states 2 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-carriers 1
full kfinalize_4096 regex:k_finalize 2 python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --e2e-carriers 1
rm -f $OUT/*.ncu-rep
ls -la $OUT; cat $OUT/status.txt
