#!/bin/bash
# Evidence for profiles/: launch list of the bench command + full captures of the two dominant kernels.
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fuse_blend -s 3 -c 1 -f -o gpurun_out/prof_fuse $CMD > gpurun_out/ncu_full.log 2>&1
python tools/rl_case.py > gpurun_out/rl_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rl_blur -s 8 -c 2 -f -o gpurun_out/prof_rl python tools/rl_case.py > gpurun_out/ncu_rl.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200; cat gpurun_out/rl_plain.log
