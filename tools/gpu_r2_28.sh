#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["frac"], l.get("clocks"), l["sustained"])
for r in l.get("layers", []): print("   ", r["layer"], r["ms"], round(r["tflops"]))
print(l["roofline"]["step_accounting_ms"])
PY
