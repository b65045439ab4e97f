#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/ae_eval_launches.csv python tools/prof_ae.py --batch 32 --profile-eval > gpurun_out/ae_eval_ncu.log 2>&1
echo "ncu exit=$?"; tail -1 gpurun_out/ae_eval_ncu.log
python tools/launch_summary.py gpurun_out/ae_eval_launches.csv | head -30 | tee gpurun_out/ae_eval_kernel_totals.txt
