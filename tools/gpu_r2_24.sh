#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused12.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 25 | cut -c1-300
