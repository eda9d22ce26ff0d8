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
ncu --set full --clock-control none --import-source on -k regex:rl_xy_kernel -s 8 -c 1 -f -o gpurun_out/prof_rl_xy python tools/rl_case.py > gpurun_out/ncu_rl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rl_z_kernel -s 8 -c 1 -f -o gpurun_out/prof_rl_z python tools/rl_case.py >> gpurun_out/ncu_rl.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rl_ -s 12 -c 12 --csv --log-file gpurun_out/rl_v3_launches.csv python tools/rl_case.py >> gpurun_out/ncu_rl.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200; cat gpurun_out/rl_plain.log
