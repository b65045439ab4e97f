#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/test_gpu_conv.log 2>&1
echo "exit=$?" >> gpurun_out/test_gpu_conv.log
grep -v "^$" gpurun_out/test_gpu_conv.log | grep -E "^E |passed|failed|^FAILED|exit" | cut -c1-300 | head -30
timeout 200 python tools/bench_last.py 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-ae --no-train-ast --layers-out gpurun_out/layers.json 2>gpurun_out/bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'whole', d['roofline']['whole_step_tflops']); print('train', d['train']['value'], d['train']['ms_per_step'])"
