#!/bin/bash
# round 2: K4 path with fp16 activations -- unit tests first, then the whole-network tests with printed errors
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mobile.py tests/test_gpu_dropin.py tests/test_gpu_attn.py -q -m gpu -p no:cacheprovider -s > gpurun_out/test_k4_f16.log 2>&1; echo "exit=$?" >> gpurun_out/test_k4_f16.log
grep -E "^FAILED|^ERROR|passed|failed|AutoEncoder 256|eval-mode relative|relative L2 vs the|gradient-norm ratio|ragged|trainer loop" gpurun_out/test_k4_f16.log | cut -c1-1800
grep -E "^E " gpurun_out/test_k4_f16.log | head -40 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit=$?" >> gpurun_out/smoke.log
tail -n 5 gpurun_out/smoke.log | cut -c1-400
