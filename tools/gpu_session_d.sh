#!/bin/bash
set -u
OUT=gpurun_out/${1:-r02d}; mkdir -p $OUT
echo "== default"; timeout 300 python tools/debug/dbg_c3.py 2>&1 | tee $OUT/dbg_default.log
echo "== TETRA_PFB=0"; TETRA_PFB=0 timeout 300 python tools/debug/dbg_c3.py 2>&1 | tee $OUT/dbg_nopfb.log
echo "== TETRA_EDGE_MODE=3"; TETRA_EDGE_MODE=3 timeout 300 python tools/debug/dbg_c3.py 2>&1 | tee $OUT/dbg_edge3.log
