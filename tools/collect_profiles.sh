#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (tracked): the files the round's evidence README quotes.
set -e
G=gpurun_out; P=profiles
cp $G/launches.csv $P/r1_launches_inference_step.csv
cp $G/train_launches.csv $P/r1_launches_train_step.csv
cp $G/ae_train_launches.csv $P/r1_launches_ae_train_step.csv
cp $G/ast_train_launches.csv $P/r1_launches_ast_train_step.csv
cp $G/adaattn_layer_launches.csv $P/r1_launches_adaattn_layer.csv
python tools/launch_summary.py $P/r1_launches_inference_step.csv > $P/r1_inference_step_kernel_totals.txt
python tools/launch_summary.py $P/r1_launches_train_step.csv > $P/r1_train_step_kernel_totals.txt
python tools/launch_summary.py $P/r1_launches_ae_train_step.csv > $P/r1_ae_train_step_kernel_totals.txt
python tools/launch_summary.py $P/r1_launches_ast_train_step.csv > $P/r1_ast_train_step_kernel_totals.txt
python tools/launch_summary.py $P/r1_launches_adaattn_layer.csv > $P/r1_adaattn_layer_kernel_totals.txt
python tools/ncu_condense.py $G/ncu_prof_raw.csv "ncu --set full --clock-control none: tools/prof_target.py 8 (K1 + one stylise pass, batch 8 at 512^2)" > $P/r1_ncu_full_summary.csv
python tools/ncu_condense.py $G/ncu_k1_raw.csv "ncu --set full: tools/bench_k1.py, adain_cached_kernel at (32,512,64,64) fp32" > $P/r1_ncu_k1_summary.csv
python tools/ncu_condense.py $G/ncu_dw_raw.csv "ncu --set full: tools/bench_dw.py --n 32 --only 240x5" > $P/r1_ncu_depthwise_summary.csv
python tools/ncu_condense.py $G/ncu_pw_raw.csv "ncu --set full: pointwise / weight-gradient GEMM launches inside the config-3 step" > $P/r1_ncu_pointwise_summary.csv
python tools/ncu_condense.py $G/ncu_attn_raw.csv "ncu --set full: tools/prof_ast.py --layer, one AdaAttN layer forward + backward at (8,128,32,32)" > $P/r1_ncu_adaattn_summary.csv
python tools/ncu_condense.py $G/ncu_last_raw.csv "ncu --set full: tools/bench_last.py, conv3x3_last_tn_kernel at (32,64,512,512)" > $P/r1_ncu_last_layer_summary.csv
python tools/ncu_condense.py $G/ncu_first_raw.csv "ncu --set full: tools/bench_first.py, conv3x3_first_tma_kernel at (32,3,512,512)" > $P/r1_ncu_first_layer_summary.csv
python tools/ncu_traffic.py $G/ncu_first_raw.csv conv3x3_first_tma conv3x3_first_tma_kernel
python tools/ncu_traffic.py $G/ncu_k1_raw.csv adain_cached_kernel adain_cached_kernel
python tools/ncu_traffic.py $G/ncu_last_raw.csv conv3x3_last_tn conv3x3_last_tn_kernel
python tools/ncu_traffic.py $G/ncu_attn_raw.csv attn_softmax_kernel attn_softmax_kernel
cp $G/layers.json $P/r1_layers_table.json
cp $G/bench_k1.log $P/r1_bench_k1.txt
cp $G/bench_dw.log $P/r1_bench_depthwise_kernels.txt
cp $G/bench_last.log $P/r1_bench_last_layer.txt
cp $G/bench_first.log $P/r1_bench_first_layer.txt
cp $G/bw_probe.log $P/r1_bw_probe.txt
cp $G/conv_role_breakdown.txt $P/r1_conv_role_breakdown.txt
cp $G/smoke.log $P/r1_smoke.log
tail -n 3 $G/test_gpu_all.log > $P/r1_pytest_gpu.log
cp $G/gpu.txt $P/r1_gpu.txt
head -1 $G/bench.log > $P/r1_bench_1gpu.json
head -1 $G/bench_ref.log > $P/r1_bench_reference_arm.json
ls -la $P | wc -l; du -sh $P
