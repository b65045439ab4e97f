#!/bin/bash
mkdir -p gpurun_out
L="enc2 enc3 enc4 dec7 enc6"
for pr in 0 1; do
for cfg in "0 0" "200 0" "500 0" "1000 0" "500 200" "1000 500"; do
  set -- $cfg
  echo "== AST_CONV_PAIR=$pr AST_CONV_EPI_SLEEP=$1 AST_CONV_PROD_SLEEP=$2"
  AST_CONV_PAIR=$pr AST_CONV_EPI_SLEEP=$1 AST_CONV_PROD_SLEEP=$2 timeout 120 python tools/bench_conv.py $L
done; done > gpurun_out/bench_conv_sleep.txt 2>&1
cat gpurun_out/bench_conv_sleep.txt
