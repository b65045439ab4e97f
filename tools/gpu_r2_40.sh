#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ae_small_batch.py 2>&1 | tail -6 | cut -c1-200 | tee gpurun_out/ae_small_batch.txt
