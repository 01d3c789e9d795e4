#!/bin/bash
set -u
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -s -k "stft or waterfall or spectrum" > $OUT/pytest_stft.log 2>&1; echo "pytest stft rc=$?" | tee -a $OUT/status.txt
grep -v "^$" $OUT/pytest_stft.log | tail -15
for P in 1 0; do
TETRA_STFT_PAIR=$P TETRA_CONFIGS=5 timeout 600 python tools/bench_configs.py > $OUT/configs5_pair$P.json 2> $OUT/configs5_$P.err; echo "configs5 pair=$P rc=$?" | tee -a $OUT/status.txt
cat $OUT/configs5_pair$P.json
done
TETRA_CONFIGS=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stft4096_db2 --launch-skip 60 -c 1 -o $OUT/stft_pair -f python tools/bench_configs.py > $OUT/ncu_stft.log 2>&1; echo "ncu rc=$?" | tee -a $OUT/status.txt
ncu -i $OUT/stft_pair.ncu-rep --page details > $OUT/stft_pair_details.txt 2>/dev/null
grep -E "^\s+(Duration|DRAM Throughput|Issue Slots Busy|Registers Per Thread|Achieved Occupancy|Executed Ipc Active|Memory Throughput)" $OUT/stft_pair_details.txt
