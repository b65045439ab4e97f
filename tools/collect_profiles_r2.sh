#!/bin/bash
# gpurun_out/ (scratch, written by tools/gpu_r2_final.sh) -> profiles/ (tracked): the files profiles/README.md quotes.
set -e
G=gpurun_out; P=profiles
cp $G/launches.csv $P/r2_launches_inference_step.csv
python tools/launch_summary.py $P/r2_launches_inference_step.csv > $P/r2_inference_step_kernel_totals.txt
cp $G/ae_train_launches.csv $P/r2_launches_ae_train_step.csv
python tools/launch_summary.py $P/r2_launches_ae_train_step.csv > $P/r2_ae_train_step_kernel_totals.txt
cp $G/train_launches.csv $P/r2_launches_train_step.csv
python tools/launch_summary.py $P/r2_launches_train_step.csv > $P/r2_train_step_kernel_totals.txt
python tools/ncu_condense.py $G/r2_step_full_raw.csv "ncu --set full --clock-control none: tools/prof_step.py 32 (one stylise pass at the bench shape, batch 32 at 512^2)" > $P/r2_ncu_step_b32_summary.csv
python tools/ncu_family_traffic.py $G/r2_step_full_raw.csv conv3x3_pair_kernel "r2_step_full_raw.csv (ncu --set full --clock-control none, tools/prof_step.py 32)" conv3x3_pair_kernel conv3x3_fold_pair_kernel conv12_fused_pair_kernel
python tools/ncu_traffic.py $G/r2_step_full_raw.csv conv3x3_last_tn conv3x3_last_tn_kernel
python tools/ncu_traffic.py $G/r2_step_full_raw.csv conv12_fused_pair_kernel conv12_fused_pair_kernel
cp $G/layers.json $P/r2_layers_table.json
cp $G/conv12_fused_timing.txt $P/r2_conv12_fused_timing.txt
cp $G/ae_small_batch.txt $P/r2_ae_step_vs_batch.txt
cp $G/bench_dw.txt $P/r2_bench_depthwise_kernels.txt
cp $G/bench_wgrad.txt $P/r2_bench_wgrad_native.txt
cp $G/bench_pw.txt $P/r2_bench_pointwise_kernels.txt
cp $G/ae_b32.log $P/r2_ae_step_b32.json
cp $G/smoke.log $P/r2_smoke.log
tail -n 3 $G/test_gpu_all.log > $P/r2_pytest_gpu.log
cp $G/gpu.txt $P/r2_gpu.txt
head -1 $G/bench.log > $P/r2_bench_1gpu.json
head -1 $G/bench_ref.log > $P/r2_bench_reference_arm.json
ls $P | wc -l; du -sh $P
