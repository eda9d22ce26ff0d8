#!/bin/bash
# first GPU call: parity tests + pipe micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
./tools/ubench_pipes > gpurun_out/ubench_pipes.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/ubench_pipes.txt gpurun_out/pytest_gpu.txt
