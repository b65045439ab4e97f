#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider -k "ragged or too_small" -s > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
grep -v "^$" gpurun_out/test_gpu_mobile.log | grep -E "^E |passed|failed|^tests|exit|ragged" | cut -c1-600 | head -30
