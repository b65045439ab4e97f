#!/bin/bash
# BatchNorm counter inside ast_bn_finalize + one zero fill per block: all GPU tests, config-3 step vs batch size
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/test_gpu_all.log
timeout 300 python tools/ae_small_batch.py 2>&1 | tail -4 | cut -c1-120 | tee gpurun_out/ae_small_batch.txt
