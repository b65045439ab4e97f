#!/bin/bash
mkdir -p gpurun_out
for pr in 0 1; do
AST_CONV_PAIR=$pr AST_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py enc2 enc3 enc4 dec5 dec7 enc6 2>&1 | grep "conv dbg" | awk 'NR%13==0' | cut -c1-400
done > gpurun_out/role_breakdown_pair.txt
cat gpurun_out/role_breakdown_pair.txt
