#!/bin/bash
# ncu capture of the fused kernel (run only after the plain command exited 0)
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fuse_blend -s 3 -c 1 -f -o gpurun_out/prof_fuse $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/plain.log
tail -5 gpurun_out/ncu_full.log
