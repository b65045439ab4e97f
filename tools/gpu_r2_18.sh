#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 4 | cut -c1-300
for wa in 0 1; do
echo "== AST_CONV_WIDEA=$wa"
AST_CONV_WIDEA=$wa timeout 120 python tools/bench_conv.py enc2 enc3 enc4 enc5 enc6 enc9 dec1 dec5 dec7
done 2>&1 | tee gpurun_out/bench_conv_widea.txt
AST_CONV_DEBUG=1 timeout 120 python tools/bench_conv.py enc2 enc3 enc4 dec5 dec7 enc6 2>&1 | grep "conv dbg" | awk 'NR%13==0' | cut -c1-400 | tee gpurun_out/role_breakdown_pair.txt
