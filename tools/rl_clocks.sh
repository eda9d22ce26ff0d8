#!/bin/bash
# run tools/rl_case.py (longer loop) while sampling SM clock / power / throttle reasons
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown,temperature.gpu --format=csv,noheader -lms 200 > gpurun_out/rl_smi.txt &
SMI=$!
RL_ITERS=40 python tools/rl_case.py | tail -1
kill $SMI
sort gpurun_out/rl_smi.txt | uniq -c | sort -rn | head -8
