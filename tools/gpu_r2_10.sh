#!/bin/bash
# CTA-pair conv kernel: parity tests, then A/B timing of the step's layers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_pair.log 2>&1; echo "exit=$?" >> gpurun_out/test_pair.log
tail -n 25 gpurun_out/test_pair.log | cut -c1-300
L="enc2 enc3 enc4 enc5 enc6 enc9 dec1 dec3 dec5 dec7"
for pr in 0 1; do
  echo "== AST_CONV_PAIR=$pr"
  AST_CONV_PAIR=$pr timeout 120 python tools/bench_conv.py $L
done > gpurun_out/bench_conv_pair.txt 2>&1
cat gpurun_out/bench_conv_pair.txt
