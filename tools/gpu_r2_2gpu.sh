#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_ae_check.py > gpurun_out/dp_ae_check_$N.log 2>&1; echo "exit=$?"
tail -n 3 gpurun_out/dp_check_$N.log gpurun_out/dp_ae_check_$N.log | cut -c1-600
python - <<PY
import json
l = json.loads([x for x in open("gpurun_out/bench_${N}gpu.log").read().strip().splitlines() if x.startswith("{")][-1])
print(l["n_gpus"], l["value"], l["ms_per_step"], "e2e", l["e2e"]["value"], "train", l["train"]["value"], l["train"].get("allreduce_bytes_per_step"), "ae", l["train_ae"]["value"], l["train_ae"].get("mode"))
PY
