#!/bin/bash
# RL pair-kernel tile configurations (library built with MVF_NVCC_EXTRA=-DMVF_RL_TUNE) vs the scalar kernel
export MVF_NVCC_EXTRA=-DMVF_RL_TUNE
echo -n "v1 (scalar FFMA): "; MVF_RL_KERNEL=v1 python tools/rl_case.py | tail -1
for c in 0 1 2 3 4; do echo -n "pair cfg $c: "; MVF_RL_CFG=$c python tools/rl_case.py | tail -1; done
