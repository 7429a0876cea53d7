#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests + smoke; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
