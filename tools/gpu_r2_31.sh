#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_attn.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -n 5 | cut -c1-300
for alt in "" tools/ubench/_alt/libpw_orig.so; do
  PW_ALT=$alt timeout 120 python tools/prof_pw.py 40 240 1 1
  PW_ALT=$alt timeout 120 python tools/prof_pw.py 40 160 0 0
  PW_ALT=$alt timeout 120 python tools/prof_pw.py 160 40 0 0
done
timeout 300 python tools/bench_pw.py 2>&1 | tail -n 60 > gpurun_out/bench_pw.txt; tail -1 gpurun_out/bench_pw.txt
timeout 600 python tools/prof_ae.py --batch 32 --steps 5 2>&1 | tail -n 2
PW_ALT="" timeout 300 ncu --set full --clock-control none -k regex:pw_conv -s 3 -c 1 -o /tmp/pw_0 -f python tools/prof_pw.py 40 240 1 1 > gpurun_out/ncu_pw_0.log 2>&1
ncu -i /tmp/pw_0.ncu-rep --page raw --csv > gpurun_out/pw_raw_0.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/pw_raw_0.csv "variant 0" > gpurun_out/pw_summary_0.csv
