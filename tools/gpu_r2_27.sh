#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300 | tee gpurun_out/test_gpu_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l.get("clocks"), l["sustained"])
for r in l.get("layers", [])[:2]: print("   ", r)
print(l["train"]["value"], l["train_ae"]["value"], l["train_ast"]["value"])
PY
