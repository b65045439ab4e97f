#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_attn.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -n 4 | cut -c1-300
for r in 0 1; do
  AST_PW_RESIDENT=$r timeout 120 python tools/prof_pw.py 16 96 0 0 | sed "s/^/resident $r: /"
  AST_PW_RESIDENT=$r timeout 120 python tools/prof_pw.py 40 240 1 1 | sed "s/^/resident $r: /"
  AST_PW_RESIDENT=$r timeout 120 python tools/prof_pw.py 160 40 0 0 | sed "s/^/resident $r: /"
done
timeout 300 python tools/bench_pw.py 2>&1 | tail -n 60 > gpurun_out/bench_pw.txt; tail -1 gpurun_out/bench_pw.txt
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -n 1
