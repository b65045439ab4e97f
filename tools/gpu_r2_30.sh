#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_pw.py 2>&1 | tail -n 60 | tee gpurun_out/bench_pw.txt
