#!/bin/bash
# full pass: tests (one process per file), smoke, bench (+ per-layer table + config-2 step), ncu
mkdir -p gpurun_out
for f in test_gpu_train test_gpu_conv test_gpu_pipeline test_gpu_adain test_gpu_losses; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout=600 > gpurun_out/$f.log 2>&1
  echo "exit=$?" >> gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "exit=$?" >> gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/plain_bench.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit=$?" >> gpurun_out/ncu_launches.log
  timeout 300 python tools/prof_target.py 8 > gpurun_out/plain_prof.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"adain_cached|conv3x3|native_" -c 36 \
      -o gpurun_out/prof -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?" >> gpurun_out/ncu_full.log
fi
tail -n 3 gpurun_out/*.log
