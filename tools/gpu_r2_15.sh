#!/bin/bash
mkdir -p gpurun_out
for tg in 1 2 4; do
echo "== AST_CONV_TG=$tg"
AST_CONV_TG=$tg timeout 120 python tools/bench_conv.py enc2 enc3 enc4 enc5 enc6 enc8 dec1 dec3 dec5 dec7
done > gpurun_out/bench_conv_tg.txt 2>&1
cat gpurun_out/bench_conv_tg.txt
AST_CONV_TG=4 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_pair.log 2>&1; echo "exit=$?" >> gpurun_out/test_pair.log
tail -n 6 gpurun_out/test_pair.log | cut -c1-300
