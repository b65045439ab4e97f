#!/bin/bash
# First GPU pass: smoke, diagnostics, parity tests (one process per file so that a faulting kernel
# cannot poison the others), a short bench and the ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag" > gpurun_out/diag.log
for args in "1 16 16 64 64 0 0 1" "1 16 16 64 64 0 0 64" "1 32 32 128 256 0 0 1" "2 24 40 64 128 1 0 1" "1 16 32 256 512 2 1 1" "4 64 64 128 128 0 1 1"; do
  echo "-- diag_conv $args" >> gpurun_out/diag.log
  timeout 180 python tools/diag_conv.py $args >> gpurun_out/diag.log 2>&1
  echo "exit=$?" >> gpurun_out/diag.log
done
for f in test_gpu_adain test_gpu_losses test_gpu_conv test_gpu_pipeline; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 > gpurun_out/$f.log 2>&1
  echo "exit=$?" >> gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "exit=$?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/*.log
