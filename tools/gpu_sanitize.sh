#!/bin/bash
# compute-sanitizer over the small parity cases of every kernel family (memcheck) and over the fused path's shared-memory
# pipeline (racecheck on the smoke case). Usage (under gpurun): bash tools/gpu_sanitize.sh [tag]
set -u
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL='tiny or rate_ or gui or unaligned or one_symbol or mixed or outside_fused or u8_ingest or u8_with or survey or spectrum or shapes or peer_memory or allgather or transport or edge or crafted or bursts or chunked or longer_than or as_many_items'
timeout 2400 compute-sanitizer --tool memcheck --padding 4096 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/status.txt
tail -6 $OUT/memcheck.log
timeout 1800 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/status.txt
tail -6 $OUT/racecheck.log
# shared-memory pipelines of every mode of the fused kernel (float / freq_offset / bytes / bytes + freq_offset / wideband), the
# finalize kernel's fused front end and the exchange kernels
RSEL='gui or min_fast or u8_ingest_fused_path or u8_with_offsets or config3_wideband_channels or crafted or ragged or peer_memory or allgather_fused or chunked_host'
timeout 1800 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -q -x -k "$RSEL" >> $OUT/racecheck.log 2>&1; echo "racecheck-modes rc=$?" | tee -a $OUT/status.txt
tail -6 $OUT/racecheck.log
