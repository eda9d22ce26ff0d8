#!/bin/bash
set -e
CMD="python tools/rl_case.py"
$CMD > gpurun_out/rl_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rl_blur -s 8 -c 2 -f -o gpurun_out/prof_rl $CMD > gpurun_out/ncu_rl.log 2>&1
cat gpurun_out/rl_plain.log; tail -3 gpurun_out/ncu_rl.log
