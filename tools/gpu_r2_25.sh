#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused12.py tests/test_gpu_pipeline.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 8 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fused12.log 2> gpurun_out/bench_fused12.err; echo "bench rc=$?"
AST_CONV12_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_unfused12.log 2> gpurun_out/bench_unfused12.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench_fused12", "bench_unfused12"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l.get("clocks"))
        for r in l.get("layers", [])[:3]: print("   ", r)
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_fused12.err
