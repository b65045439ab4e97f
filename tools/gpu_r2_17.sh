#!/bin/bash
mkdir -p gpurun_out
L="enc2 enc3 enc4 dec7 enc6"
for f in 0 2 4 8; do
  echo "== AST_CONV_PF=$f"
  AST_CONV_PF=$f timeout 120 python tools/bench_conv.py $L
done 2>&1 | tee gpurun_out/bench_conv_pf.txt
