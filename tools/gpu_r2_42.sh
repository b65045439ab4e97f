#!/bin/bash
# launch list of the config-2 step with K2wn
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/train_launches.csv python tools/prof_train.py --profile > gpurun_out/train_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/train_launches.csv > gpurun_out/train_step_kernel_totals.txt
head -45 gpurun_out/train_step_kernel_totals.txt
