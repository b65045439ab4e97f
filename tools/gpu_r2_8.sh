#!/bin/bash
# round 2 re-entry baseline: the full GPU suite, smoke, 1-GPU bench with the per-layer table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
tail -n 30 gpurun_out/test_gpu_all.log | cut -c1-300
tail -n 4 gpurun_out/smoke.log | cut -c1-400
tail -c 800 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().split(chr(10))[0])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'))
    print('sustained', d.get('sustained'))
    r=d['roofline']; print('roofline', {k:v for k,v in r.items() if k!='layers'})
    for l in d.get('layers',[]): print(l)
    print('train', d['train'].get('value'), d['train'].get('mode'), 'ae', d['train_ae'].get('value'), d['train_ae'].get('mode'), 'ast', d.get('train_ast',{}).get('value'))
    print('edge', d['edge_layers'])
    print('adain', d['adain_roofline'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
