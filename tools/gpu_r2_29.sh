#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_attn.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -n 5 | cut -c1-300
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -n 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/ae_train_launches.csv python tools/prof_ae.py --batch 32 --profile > gpurun_out/ae_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/ae_train_launches.csv | head -40 | tee gpurun_out/ae_train_step_kernel_totals.txt
