#!/bin/bash
# single-GPU bench (both arms) exactly as the driver runs it
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit=$?" >> gpurun_out/bench_ref.log
tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.log
