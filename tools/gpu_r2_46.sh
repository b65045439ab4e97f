#!/bin/bash
# ncu --set full of K2wn, one launch per decoder layer at the config-2 shapes
mkdir -p gpurun_out
python tools/prof_wgrad.py > gpurun_out/prof_wgrad_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad3x3 -o /tmp/r2_wgrad -f python tools/prof_wgrad.py > gpurun_out/prof_wgrad_ncu.log 2>&1
echo "ncu exit=$?"
ncu -i /tmp/r2_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_wgrad_raw.csv 2> gpurun_out/ncu_export.err
python tools/ncu_condense.py gpurun_out/r2_wgrad_raw.csv "ncu --set full --clock-control none: tools/prof_wgrad.py (K2wn, the nine decoder layers at batch 8, 256x256 image)" > gpurun_out/r2_ncu_wgrad_native_summary.csv
cut -c1-400 gpurun_out/r2_ncu_wgrad_native_summary.csv | head -14
