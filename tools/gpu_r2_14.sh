#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_pair.log 2>&1; echo "exit=$?" >> gpurun_out/test_pair.log
tail -n 12 gpurun_out/test_pair.log | cut -c1-300
for pr in 0 1; do
echo "== AST_CONV_PAIR=$pr"
AST_CONV_PAIR=$pr timeout 120 python tools/bench_conv.py
done > gpurun_out/bench_conv_pair.txt 2>&1
cat gpurun_out/bench_conv_pair.txt
AST_CONV_PAIR=1 AST_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py enc2 enc3 enc4 dec5 dec7 2>&1 | grep "conv dbg" | awk 'NR%13==0' | cut -c1-400 > gpurun_out/role_breakdown_pair.txt
cat gpurun_out/role_breakdown_pair.txt
