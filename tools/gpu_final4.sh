#!/bin/bash
# Last refresh of the round: full GPU test suite, smoke, the bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/test_gpu_all.log gpurun_out/smoke.log | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().split(chr(10))[0])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['whole_step_tflops'], 'train', d['train']['value'], 'ae', d['train_ae']['value'], 'ast', d['train_ast']['value'])
print('edge', d['edge_layers']['first']['frac'], d['edge_layers']['last']['frac'])"
