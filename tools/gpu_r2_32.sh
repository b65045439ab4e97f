#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_attn.py tests/test_gpu_dropin.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -n 5 | cut -c1-300
timeout 300 python tools/bench_pw.py 2>&1 | tail -n 60 > gpurun_out/bench_pw.txt; tail -1 gpurun_out/bench_pw.txt
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -n 2
