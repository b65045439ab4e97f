#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_mobile.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_mobile.log
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 > gpurun_out/ae_b32.log 2>&1; echo "exit=$?" >> gpurun_out/ae_b32.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/ae_train_launches.csv python tools/prof_ae.py --batch 32 --profile > gpurun_out/ae_ncu.log 2>&1
echo "ncu exit=$?" >> gpurun_out/ae_ncu.log
for f in gpurun_out/test_gpu_mobile.log gpurun_out/ae_b32.log gpurun_out/ae_ncu.log; do tail -n 3 $f; done
