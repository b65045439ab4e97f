#!/bin/bash
# AST / AdaAttN: all GPU tests + smoke, the AST step timing, ncu launch lists of one step and of one layer fwd+bwd
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu_all.log 2>&1; echo "exit=$?" >> gpurun_out/test_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
timeout 600 python tools/prof_ast.py > gpurun_out/ast_step.log 2>&1; echo "exit=$?" >> gpurun_out/ast_step.log
timeout 300 python tools/prof_ast.py --profile > gpurun_out/ast_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/ast_train_launches.csv python tools/prof_ast.py --profile > gpurun_out/ast_ncu.log 2>&1
timeout 300 python tools/prof_ast.py --layer > gpurun_out/ast_layer_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/adaattn_layer_launches.csv python tools/prof_ast.py --layer > gpurun_out/ast_layer_ncu.log 2>&1
tail -n 4 gpurun_out/test_gpu_all.log gpurun_out/smoke.log gpurun_out/ast_step.log gpurun_out/ast_plain.log gpurun_out/ast_layer_plain.log | cut -c1-1500
