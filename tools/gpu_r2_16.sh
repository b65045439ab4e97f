#!/bin/bash
mkdir -p gpurun_out
L="enc2 enc3 dec7 dec8 dec6"
for f in 0 1 2 4 6 7; do
  echo "== AST_CONV_DBGFLAGS=$f (1 = no operand reloads, 2 = no stores, 4 = no TMEM loads / epilogue math)"
  AST_CONV_DBGFLAGS=$f timeout 120 python tools/bench_conv.py $L
done > gpurun_out/bench_conv_elim_pair.txt 2>&1
cat gpurun_out/bench_conv_elim_pair.txt
