#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_losses.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_losses.log
grep -v "^$" gpurun_out/test_gpu_losses.log | grep -E "^E |passed|failed|^tests|exit" | cut -c1-400 | head -20
