#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3 | cut -c1-300
timeout 300 python tools/bench_pw.py > gpurun_out/bench_pw.txt 2>&1; tail -1 gpurun_out/bench_pw.txt
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -n 1 | tee gpurun_out/ae_b32.log
