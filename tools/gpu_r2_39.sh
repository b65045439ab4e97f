#!/bin/bash
timeout 300 python tools/bench_dw.py --n 32 2>&1 | tail -24 | tee gpurun_out/bench_dw.txt
