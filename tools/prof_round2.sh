#!/bin/bash
# Round-2 evidence for profiles/r02: launch list of the bench command + full captures of the dominant kernels
# (fuse_axis_kernel of configs[1]; rl_zz_kernel / rl_xy_kernel<CORRECT> / rl_xy_kernel<PLAIN> of configs[3]).
# Every ncu pass follows a plain run of the same command that exited 0.
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
RL="python tools/rl_case.py"
export RL_DISTINCT=1 RL_ITERS=1
$CMD > gpurun_out/r02_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
python tools/fuse_axis_case.py f32 > gpurun_out/r02_fuse_axis_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fuse_axis_kernel -s 3 -c 1 -f -o gpurun_out/r02_fuse_axis python tools/fuse_axis_case.py f32 > gpurun_out/r02_ncu_fuse.log 2>&1
$RL > gpurun_out/r02_rl_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rl_zz_kernel -s 2 -c 1 -f -o gpurun_out/r02_rl_zz $RL > gpurun_out/r02_ncu_rl.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:rl_xy_kernel<19, 4, 2>" -s 5 -c 1 -f -o gpurun_out/r02_rl_xyc $RL >> gpurun_out/r02_ncu_rl.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:rl_xy_kernel<19, 4, 0>" -s 5 -c 1 -f -o gpurun_out/r02_rl_xy $RL >> gpurun_out/r02_ncu_rl.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rl_ -s 28 -c 28 --csv --log-file gpurun_out/r02_rl_v4_launches.csv $RL >> gpurun_out/r02_ncu_rl.log 2>&1
tail -1 gpurun_out/r02_plain.log | cut -c1-160; cat gpurun_out/r02_rl_plain.log gpurun_out/r02_fuse_axis_plain.log
