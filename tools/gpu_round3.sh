#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_train test_gpu_conv test_gpu_pipeline test_gpu_adain test_gpu_losses; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout=600 -x -s > gpurun_out/$f.log 2>&1
  echo "exit=$?" >> gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "exit=$?" >> gpurun_out/bench.log
tail -n 4 gpurun_out/*.log
