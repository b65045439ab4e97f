#!/bin/bash
# K2wn (native-layout 3x3 weight gradient): parity in both B-load modes, micro-benchmark, config-2 step A/B
mkdir -p gpurun_out
{
echo "== haloed B box (default)"; timeout 300 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider 2>&1 | tail -8
echo "== AST_WGRAD_SEP=1"; AST_WGRAD_SEP=1 timeout 300 python -m pytest tests/test_gpu_train.py -q -x -p no:cacheprovider -k wgrad_native 2>&1 | tail -8
} > gpurun_out/wgrad_native_tests.txt 2>&1
timeout 300 python tools/bench_wgrad.py --step > gpurun_out/bench_wgrad.txt 2>&1
AST_WGRAD_SEP=1 timeout 300 python tools/bench_wgrad.py > gpurun_out/bench_wgrad_sep.txt 2>&1
AST_WGRAD_PLANAR=1 timeout 300 python tools/bench_wgrad.py --step 2>&1 | tail -1 > gpurun_out/bench_wgrad_planar_step.txt
cat gpurun_out/wgrad_native_tests.txt; cat gpurun_out/bench_wgrad.txt; tail -3 gpurun_out/bench_wgrad_sep.txt; cat gpurun_out/bench_wgrad_planar_step.txt
