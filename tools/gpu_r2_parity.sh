#!/bin/bash
# the parity tests at the benchmarked sizes and the whole-network tests, with their printed error figures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_mobile.py tests/test_gpu_dropin.py tests/test_gpu_train.py -q -m gpu -p no:cacheprovider -s \
  -k "config4 or config5 or config1 or 256 or trainer or loss_curve or non_degenerate or full_training_step or per_batch" > gpurun_out/parity_bench_sizes.log 2>&1; echo "exit=$?" >> gpurun_out/parity_bench_sizes.log
grep -v "^$" gpurun_out/parity_bench_sizes.log | cut -c1-1500 | tail -60
