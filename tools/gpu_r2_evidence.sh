#!/bin/bash
# round 2 evidence: bench (1 GPU) + reference arm, ncu launch list of the same bench command, ncu --set full over one
# stylise pass at the bench shape (raw page CSV only: the .ncu-rep exceeds the return limit), GPU test log, smoke.
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
tail -n 3 gpurun_out/test_gpu_all.log; tail -n 3 gpurun_out/smoke.log | cut -c1-300; head -c 600 gpurun_out/bench.log; echo; head -c 600 gpurun_out/bench_ref.log; echo
ls -la gpurun_out/r2_step_full_raw.csv gpurun_out/launches.csv
