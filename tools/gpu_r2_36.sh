#!/bin/bash
for f in "" 8; do
  for sh in "160 40 0 0" "96 16 0 0" "40 240 1 1" "16 96 0 0" "320 40 0 1 128"; do
    AST_PW_DBGFLAGS=$f timeout 120 python tools/prof_pw.py $sh | sed "s/^/dbg '$f': /"
  done
done
timeout 300 python tools/bench_pw.py 2>&1 | tail -n 60 > gpurun_out/bench_pw.txt; tail -1 gpurun_out/bench_pw.txt
