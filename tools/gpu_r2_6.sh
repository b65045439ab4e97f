#!/bin/bash
# round 2: 2-GPU bench (config 4 sharded; config 2 and config 3 training with the NCCL all-reduce inside the CUDA graph)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 --no-train-ast > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "exit=$?" >> gpurun_out/bench_2gpu.log
tail -c 1200 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_2gpu.log').read().split(chr(10))[0])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'f32', d['e2e']['fp32_host_tensors']['value'])
    print('train', {k:d['train'].get(k) for k in ('value','mode','eager_steps_per_s','img_per_s','error')})
    print('train_ae', {k:d['train_ae'].get(k) for k in ('value','mode','eager_steps_per_s','img_per_s','error')})
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_2gpu.log').read()[-1500:])
PY
