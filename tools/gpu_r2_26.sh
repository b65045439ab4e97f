#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused12.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3 | cut -c1-300
{
for v in 0 1 2 3; do SUSTAINED=1 AST_CONV12_V=$v timeout 120 python tools/bench_conv12.py 2>&1 | tail -1; done
for f in 2 6 64 128 256 454; do AST_CONV_DBGFLAGS=$f timeout 120 python tools/bench_conv12.py 2>&1 | tail -1; done
AST_CONV_DEBUG=1 timeout 120 python tools/bench_conv12.py 2>&1 | grep "conv12 dbg" | tail -1
} | tee gpurun_out/bench_conv12_elim.txt
