#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_k1.py > gpurun_out/bench_k1.log 2>&1; echo "exit=$?" >> gpurun_out/bench_k1.log
timeout 600 python -m pytest tests/test_gpu_adain.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_adain.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_adain.log
cat gpurun_out/bench_k1.log; tail -n 5 gpurun_out/test_gpu_adain.log
