#!/bin/bash
mkdir -p gpurun_out
for F in 3 5 6 7; do
timeout 400 python bench.py --steps 8 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline --skip-roofline --frames-per-step $F > gpurun_out/r2_c25_f$F.json 2> gpurun_out/r2_c25_f$F.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c25_f$F.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("strict fps$F:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c25_f$F.err").read()[-1500:])
PY
done
