#!/bin/bash
# halo-ring zeroing + frozen-weight pack cache: all GPU tests, config-2 step, config-3 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
tail -5 gpurun_out/test_gpu_all.log
timeout 300 python tools/bench_wgrad.py --step 2>&1 | tail -2
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -1
