#!/bin/bash
# Round-end evidence run: tests, smoke, bench (+layers, +train), per-role wait breakdown, ncu launch lists
# (inference + training step) and one --set full capture of the hot kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_train test_gpu_conv test_gpu_pipeline test_gpu_adain test_gpu_losses; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout=600 > gpurun_out/$f.log 2>&1
  echo "exit=$?" >> gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "exit=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit=$?" >> gpurun_out/bench_ref.log
AST_CONV_DEBUG=1 timeout 300 python tools/dbg_layers.py 32 2>&1 | grep "conv dbg" | awk "NR%2==0" > gpurun_out/conv_role_breakdown.txt
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> gpurun_out/ncu_launches.log
timeout 200 python tools/prof_train.py > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv \
    --log-file gpurun_out/train_launches.csv python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python tools/prof_target.py 8 > gpurun_out/plain_prof.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"adain_cached|conv3x3|native_" -c 36 \
    -o gpurun_out/prof -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?" >> gpurun_out/ncu_full.log
tail -n 3 gpurun_out/*.log | cut -c1-400
