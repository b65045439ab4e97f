#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_dw.py > gpurun_out/bench_dw.log 2>&1; echo "exit=$?" >> gpurun_out/bench_dw.log
timeout 300 python tools/bench_dw.py --n 32 --only 240x5 >> gpurun_out/bench_dw.log 2>&1
timeout 300 python tools/bench_dw.py --n 32 --only 160x5 >> gpurun_out/bench_dw.log 2>&1
timeout 300 python tools/bench_dw.py --n 32 --only 144x3 >> gpurun_out/bench_dw.log 2>&1
timeout 600 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
cat gpurun_out/bench_dw.log; tail -n 3 gpurun_out/test_gpu_mobile.log
