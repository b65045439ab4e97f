#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_attn.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
grep -v "^$" gpurun_out/test_gpu_mobile.log | grep -E "^E |passed|failed|^FAILED|exit" | cut -c1-300 | head -30
timeout 600 python tools/prof_ast.py 2>&1 | tail -1 | cut -c1-700
timeout 300 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -3 | cut -c1-600
