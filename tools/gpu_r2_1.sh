#!/bin/bash
# round 2, call 1: operand-format probe, the full GPU suite, the new benchmark-size parity tests with their printed
# errors, smoke, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 60 tools/ubench/mixed_fmt > gpurun_out/mixed_fmt.txt 2>&1; echo "exit=$?" >> gpurun_out/mixed_fmt.txt
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 600 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_pipeline.py tests/test_gpu_dropin.py -q -m gpu -p no:cacheprovider -s \
   -k "256 or config4 or config5 or trainer" > gpurun_out/test_gpu_big.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_big.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
cat gpurun_out/mixed_fmt.txt
tail -n 15 gpurun_out/test_gpu_all.log | cut -c1-400
grep -E "AutoEncoder 256|cfg4|cfg5|trainer loop|passed|failed" gpurun_out/test_gpu_big.log | cut -c1-1500
tail -n 4 gpurun_out/smoke.log | cut -c1-400
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().split(chr(10))[0])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'f32', d['e2e']['fp32_host_tensors']['value'], 'match', d['e2e'].get('equals_quantised_fp32_leg'))
    print('sustained', d.get('sustained'))
    r=d['roofline']; print('roofline', r['frac'], r['frac_executed'], r['share_of_step'], r['step_accounting_ms'])
    for l in d['layers']: print(l)
    print('train', d['train'].get('value'), d['train'].get('mode'), 'ae', d['train_ae'].get('value'), d['train_ae'].get('mode'), 'ast', d.get('train_ast',{}).get('value'))
    print('edge', d['edge_layers'])
    print('adain', d['adain_roofline']['frac'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
