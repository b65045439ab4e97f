#!/bin/bash
# Short round-end refresh after the first/last-layer kernel work: tests, smoke, bench, reference arm, the two edge-layer
# micro-benchmarks, the inference launch list and --set full captures of the two edge-layer kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit=$?" >> gpurun_out/bench_ref.log
timeout 300 python tools/bench_last.py > gpurun_out/bench_last.log 2>&1
timeout 300 python tools/bench_first.py > gpurun_out/bench_first.log 2>&1
AST_FIRST_NO_TMA=1 timeout 300 python tools/bench_first.py >> gpurun_out/bench_first.log 2>&1
AST_CONV_DEBUG=1 timeout 100 python tools/bench_first.py --reps 2 2>&1 | grep "conv dbg" | tail -1 >> gpurun_out/bench_first.log
timeout 200 python tools/ubench/bw_probe.py > gpurun_out/bw_probe.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-train-ae --no-train-ast > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-train-ae --no-train-ast > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> gpurun_out/ncu_launches.log
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_last_tn -c 2 -s 3 -o gpurun_out/last -f python tools/bench_last.py --reps 2 > gpurun_out/ncu_last.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_first_tma -c 2 -s 3 -o gpurun_out/first -f python tools/bench_first.py --reps 2 > gpurun_out/ncu_first.log 2>&1
for r in last first; do
  if [ -f gpurun_out/$r.ncu-rep ]; then
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/ncu_${r}_raw.csv 2>/dev/null
    rm -f gpurun_out/$r.ncu-rep
  fi
done
tail -n 3 gpurun_out/test_gpu_all.log gpurun_out/smoke.log gpurun_out/bench_last.log gpurun_out/bench_first.log gpurun_out/bw_probe.log gpurun_out/ncu_launches.log | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().split(chr(10))[0])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['whole_step_tflops'], 'train', d['train']['value'], 'ae', d['train_ae']['value'], 'ast', d['train_ast']['value'])"
