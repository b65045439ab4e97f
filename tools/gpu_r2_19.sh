#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 8 | cut -c1-300
(echo "== per-lane sector stores (AST_FIRST_NO_TMA_STORE=1)"; AST_FIRST_NO_TMA_STORE=1 timeout 120 python tools/bench_first.py; echo "== TMA-store epilogue"; timeout 120 python tools/bench_first.py) 2>&1 | tee gpurun_out/bench_first.txt
