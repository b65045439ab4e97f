#!/bin/bash
# AdaAttN / AST parity tests on one B200
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_attn.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_attn.log
grep -v "^$" gpurun_out/test_gpu_attn.log | grep -E "^E |passed|failed|^tests|^FAILED|exit" | cut -c1-300 | head -60
