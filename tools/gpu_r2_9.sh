#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/ubench/mma_pattern > gpurun_out/mma_pattern.txt 2>&1
cat gpurun_out/mma_pattern.txt
bash tools/gpu_r2_3.sh
AST_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py enc2 enc3 enc4 dec5 dec6 dec7 dec8 2>&1 | grep "conv dbg" | awk 'NR%13==0' | cut -c1-400 > gpurun_out/role_breakdown.txt
cat gpurun_out/role_breakdown.txt
