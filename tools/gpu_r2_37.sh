#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused12.py tests/test_gpu_conv.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 3 | cut -c1-300
for lib in "" tools/ubench/_alt/libast_before.so "" tools/ubench/_alt/libast_before.so; do
  AST_B200_LIB=$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-train-ae --no-train-ast --no-cpu-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench_ab.err
  python - "$lib" <<'PY'
import json, sys
l = json.loads(open("gpurun_out/bench_ab.log").read().strip().splitlines()[-1])
print("lib", sys.argv[1] or "new", "value", round(l["value"], 1), "ms", round(l["ms_per_step"], 3), "sustained", round(l["sustained"]["ms_per_step"], 3), l["sustained"]["sm_mhz_median"],
      {r["layer"]: round(r["ms"], 3) for r in l["layers"] if r["layer"] in ("enc_conv12", "enc_conv3", "dec_conv8", "dec_conv7", "dec_conv9", "enc_conv6")})
PY
done
