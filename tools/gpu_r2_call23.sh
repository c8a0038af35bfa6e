#!/bin/bash
mkdir -p gpurun_out
for CFG in "0,0" "0,32" "140,8" "132,16"; do
MEGA_B200_PIPE_SMS=$CFG timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline --skip-roofline --pipeline 1 > gpurun_out/r2_c23_pipe.json 2> gpurun_out/r2_c23_pipe.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c23_pipe.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("PIPE_SMS $CFG:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c23_pipe.err").read()[-1500:])
PY
done
