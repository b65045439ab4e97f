#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adain.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300
timeout 120 python tools/bench_k1.py 2>&1 | tee gpurun_out/bench_k1.txt
