#!/bin/bash
# round 2 final evidence: GPU tests, smoke, bench (1 GPU) + reference arm, ncu launch list of the bench command, ncu --set full
# over one stylise pass at the bench shape (raw page CSV), launch lists of the config-2 and config-3 steps, the fused
# conv1_1+conv1_2 kernel's micro-benchmark / role breakdown / elimination runs, the pointwise-conv table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "exit=$?" >> gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-train --no-train-ae --no-train-ast --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/prof_step.py 32 > gpurun_out/prof_step_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/r2_step_full -f python tools/prof_step.py 32 > gpurun_out/prof_step_ncu.log 2>&1
echo "ncu exit=$?"
ncu -i /tmp/r2_step_full.ncu-rep --page raw --csv > gpurun_out/r2_step_full_raw.csv 2> gpurun_out/ncu_export.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/ae_train_launches.csv python tools/prof_ae.py --batch 32 --profile > gpurun_out/ae_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/train_launches.csv python tools/prof_train.py --profile > gpurun_out/train_ncu.log 2>&1
# (the wait counters / elimination runs of the fused kernel need a build with AST_KERNEL_DEBUG=1: tools/gpu_r2_26.sh)
for v in 0 3; do SUSTAINED=1 AST_CONV12_V=$v timeout 120 python tools/bench_conv12.py 2>&1 | tail -1; done > gpurun_out/conv12_fused_timing.txt 2>&1
timeout 300 python tools/bench_wgrad.py --step > gpurun_out/bench_wgrad.txt 2>&1
timeout 300 python tools/bench_pw.py > gpurun_out/bench_pw.txt 2>&1
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 > gpurun_out/ae_b32.log 2>&1
timeout 600 python tools/ae_small_batch.py > gpurun_out/ae_small_batch.txt 2>&1
timeout 300 python tools/bench_dw.py --n 32 > gpurun_out/bench_dw.txt 2>&1
tail -n 3 gpurun_out/test_gpu_all.log; tail -n 3 gpurun_out/smoke.log | cut -c1-300; head -c 400 gpurun_out/bench.log; echo; head -c 300 gpurun_out/bench_ref.log; echo
ls -la gpurun_out/r2_step_full_raw.csv gpurun_out/launches.csv
