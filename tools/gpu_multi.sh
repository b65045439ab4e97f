#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "exit=$?" >> gpurun_out/dp_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_ae_check.py > gpurun_out/dp_ae_check_$N.log 2>&1; echo "exit=$?" >> gpurun_out/dp_ae_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_$N.log 2>&1; echo "exit=$?" >> gpurun_out/bench_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_$N.log 2>&1; echo "exit=$?" >> gpurun_out/bench_ref_$N.log
tail -n 4 gpurun_out/dp_check_$N.log gpurun_out/dp_ae_check_$N.log gpurun_out/bench_$N.log gpurun_out/bench_ref_$N.log | cut -c1-2500
