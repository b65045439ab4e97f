#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/ae_train_launches.csv python tools/prof_ae.py --batch 32 --profile > gpurun_out/ae_ncu.log 2>&1
echo "ncu exit=$?"
python tools/launch_summary.py gpurun_out/ae_train_launches.csv | head -40
