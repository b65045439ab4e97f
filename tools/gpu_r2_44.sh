#!/bin/bash
# K2wn elimination runs of the FIRST version of the kernel (commit 14aa2c0 + temporary AST_WGRAD_DBG flags: 1 no reds, 2 no MMAs, 4 no B loads, 8 no A loads;
# the flags were removed again with the fix they led to, e037751): kept as the record of how profiles/r2_wgrad_native_elimination.txt was made
mkdir -p gpurun_out
for f in 0 1 2 3 4 8 12 7 11; do echo "== AST_WGRAD_DBG=$f"; AST_WGRAD_DBG=$f timeout 120 python tools/bench_wgrad.py 2>&1 | grep -E "dec_conv(3|8|9)|sum over" | cut -c1-110; done 2>&1 | tee gpurun_out/wgrad_elim.txt
for st in 2 3; do echo "== stages $st"; AST_WGRAD_STAGES=$st timeout 120 python tools/bench_wgrad.py 2>&1 | grep -E "dec_conv(3|8)" | cut -c1-110; done 2>&1 | tee -a gpurun_out/wgrad_elim.txt
