#!/bin/bash
mkdir -p gpurun_out
L="enc2 enc4 enc6 dec6 dec7 dec8"
for cfg in "0 0" "100 0" "300 0" "1000 0" "300 100" "300 300" "2000 500"; do
  set -- $cfg
  echo "== AST_CONV_EPI_SLEEP=$1 AST_CONV_PROD_SLEEP=$2"
  AST_CONV_EPI_SLEEP=$1 AST_CONV_PROD_SLEEP=$2 timeout 120 python tools/bench_conv.py $L
done > gpurun_out/bench_conv_sleep.txt 2>&1
cat gpurun_out/bench_conv_sleep.txt
