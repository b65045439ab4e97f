#!/bin/bash
# K2wn with precomputed descriptor offsets: parity + timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider 2>&1 | tail -3
AST_WGRAD_SEP=1 timeout 300 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider -k wgrad_native 2>&1 | tail -2
timeout 300 python tools/bench_wgrad.py --step 2>&1 | tee gpurun_out/bench_wgrad.txt | cut -c1-120
