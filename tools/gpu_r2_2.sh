#!/bin/bash
# round 2, call 2: ncu --set full over one stylise pass at the bench shape (batch 32, 512x512); only the raw-page CSV
# travels back (the .ncu-rep of 30 launches exceeds the 64 MiB return limit)
mkdir -p gpurun_out
python tools/prof_step.py 32 > gpurun_out/prof_step_plain.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -o /tmp/r2_step_full -f python tools/prof_step.py 32 > gpurun_out/prof_step_ncu.log 2>&1
echo "ncu exit=$?"
tail -n 3 gpurun_out/prof_step_plain.log gpurun_out/prof_step_ncu.log
ncu -i /tmp/r2_step_full.ncu-rep --page raw --csv > gpurun_out/r2_step_full_raw.csv 2> gpurun_out/ncu_export.err
ls -la /tmp/r2_step_full* gpurun_out/r2_step_full_raw.csv | head
