#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider -k "autoencoder" -s > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
grep -v "^$" gpurun_out/test_gpu_mobile.log | grep -E "^E |passed|failed|^tests|exit|gradient-norm|eval-mode|contract" | cut -c1-1800 | head -30; tail -n 6 gpurun_out/smoke.log
