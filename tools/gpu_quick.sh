#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_losses.py tests/test_gpu_train.py tests/test_gpu_mobile.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_losses.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_losses.log
grep -v "^$" gpurun_out/test_gpu_losses.log | grep -E "^E |passed|failed|^tests|exit" | cut -c1-400 | head -20
timeout 300 python tools/prof_train.py > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv \
    --log-file gpurun_out/train_launches.csv python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train-ae 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['train'])"
