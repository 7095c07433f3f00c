#!/bin/bash
# scratch GPU job of the current experiment (edited per call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/fri-reinforcementlearning-c_b200/build/variants
for t in base q t qt; do echo "== loaded $t"; FRIQ_LIB=$V/tlx_$t.so python tools/timeline.py r2aj_$t 22 | grep -E "warp records|\[ 50, 60\)|\[ 80,100\)|\[100,200\)"; done > gpurun_out/r2aj_timeline.log 2>&1
cat gpurun_out/r2aj_timeline.log
