#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300
AST_CONV_TMA_STORE=1 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300
for ts in 0 1; do echo "== AST_CONV_TMA_STORE=$ts"; AST_CONV_TMA_STORE=$ts timeout 120 python tools/bench_conv.py enc3 dec2 dec6 dec8; done 2>&1 | tee gpurun_out/bench_conv_tma_store3.txt
