#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_train.py tests/test_gpu_attn.py tests/test_gpu_dropin.py -q -m gpu -p no:cacheprovider -s > gpurun_out/test_r2_7.log 2>&1; echo "exit=$?" >> gpurun_out/test_r2_7.log
grep -E "^FAILED|^ERROR|passed|failed|config-2 loss curve" gpurun_out/test_r2_7.log | cut -c1-600
grep -E "^E " gpurun_out/test_r2_7.log | head -30 | cut -c1-300
