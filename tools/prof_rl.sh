#!/bin/bash
# ncu --set full capture of one RL blur kernel (ratio mode) of tools/rl_case.py; KERN = kernel name regex
set -e
KERN=${KERN:-rl_pair}
CMD="python tools/rl_case.py"
$CMD > gpurun_out/rl_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERN -s 8 -c 1 -f -o gpurun_out/prof_rl $CMD > gpurun_out/ncu_rl.log 2>&1
cat gpurun_out/rl_plain.log; tail -3 gpurun_out/ncu_rl.log
