#!/bin/bash
# GPU check of the MobileNet-style (K4) path: parity tests, everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider "$@" > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
tail -5 gpurun_out/test_gpu_mobile.log
