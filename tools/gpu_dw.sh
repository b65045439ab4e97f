#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_dw.py > gpurun_out/bench_dw.log 2>&1; echo "exit=$?" >> gpurun_out/bench_dw.log
timeout 600 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider  > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_tiled_kernel -c 2 -o gpurun_out/dw_fwd --force-overwrite python tools/bench_dw.py --only 240x5 --reps 1 > gpurun_out/ncu_dw.log 2>&1
echo "ncu exit=$?" >> gpurun_out/ncu_dw.log
cat gpurun_out/bench_dw.log; tail -n 3 gpurun_out/test_gpu_mobile.log gpurun_out/ncu_dw.log
