#!/bin/bash
# Round-end evidence run (supersedes gpu_final.sh): tests, smoke, bench (+layers, +train, +train_ae, +train_ast),
# reference arm, micro-benchmarks, ncu launch lists (inference, config-2, config-3, AST step, AdaAttN layer) and
# --set full captures of the hot kernels (conv family + K1 + last layer; AdaAttN GEMM / softmax; depthwise; pointwise).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 3 --layers-out gpurun_out/layers.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit=$?" >> gpurun_out/bench_ref.log
timeout 300 python tools/bench_k1.py > gpurun_out/bench_k1.log 2>&1
timeout 300 python tools/bench_dw.py --n 32 > gpurun_out/bench_dw.log 2>&1
timeout 300 python tools/bench_last.py > gpurun_out/bench_last.log 2>&1
AST_CONV_DEBUG=1 timeout 300 python tools/dbg_layers.py 32 2>&1 | grep "conv dbg" | awk "NR%2==0" > gpurun_out/conv_role_breakdown.txt
# launch lists (each command first runs clean without ncu)
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-train-ae --no-train-ast > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-train-ae --no-train-ast > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> gpurun_out/ncu_launches.log
timeout 200 python tools/prof_train.py > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1000 --csv \
    --log-file gpurun_out/train_launches.csv python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
timeout 300 python tools/prof_ae.py --batch 32 --steps 5 > gpurun_out/ae_b32.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/ae_train_launches.csv python tools/prof_ae.py --batch 32 --profile > gpurun_out/ae_ncu.log 2>&1
timeout 300 python tools/prof_ast.py --profile > gpurun_out/ast_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/ast_train_launches.csv python tools/prof_ast.py --profile > gpurun_out/ast_ncu.log 2>&1
timeout 300 python tools/prof_ast.py --layer > gpurun_out/ast_layer_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/adaattn_layer_launches.csv python tools/prof_ast.py --layer > gpurun_out/ast_layer_ncu.log 2>&1
# full captures
timeout 300 python tools/prof_target.py 8 > gpurun_out/plain_prof.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:"adain_cached|conv3x3|native_" -c 36 \
    -o gpurun_out/prof -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?" >> gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none -k regex:adain_cached_kernel -c 2 -o gpurun_out/k1 -f python tools/bench_k1.py > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_last_tn -c 2 -s 3 -o gpurun_out/last -f python tools/bench_last.py --reps 2 > gpurun_out/ncu_last.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"bgemm_tc|attn_|split3" -c 24 -o gpurun_out/attn -f python tools/prof_ast.py --layer > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"dw_tiled|dw_wgrad_tiled" -c 6 -o gpurun_out/dw -f python tools/bench_dw.py --n 32 --only 240x5 --reps 1 > gpurun_out/ncu_dw.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"pw_conv_tc|pw_wgrad_tc" -c 8 -s 60 -o gpurun_out/pw -f python tools/prof_ae.py --batch 32 --profile > gpurun_out/ncu_pw.log 2>&1
# keep the raw-metric CSV of every capture, drop the (large) reports: gpurun_out/ travels back only below 64 MiB
for r in prof k1 last attn dw pw; do
  if [ -f gpurun_out/$r.ncu-rep ]; then
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/ncu_${r}_raw.csv 2>/dev/null
    rm -f gpurun_out/$r.ncu-rep
  fi
done
du -sh gpurun_out
tail -n 3 gpurun_out/test_gpu_all.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/bench_last.log gpurun_out/ncu_full.log gpurun_out/ncu_attn.log gpurun_out/ncu_last.log | cut -c1-400
