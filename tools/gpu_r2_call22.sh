#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_train_ops_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -k "pipelined or wavefront_step" > gpurun_out/r2_c22_pipe.log 2>&1
echo "pipelined pytest rc=$?" >> gpurun_out/r2_c22_pipe.log; tail -25 gpurun_out/r2_c22_pipe.log
for P in fp32x3 f16; do
timeout 400 python bench.py --steps 10 --warmup 3 --precision $P --no-parity --skip-cpu-baseline --skip-roofline --pipeline 1 > gpurun_out/r2_c22_bench_${P}_pipe.json 2> gpurun_out/r2_c22_bench_${P}_pipe.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c22_bench_${P}_pipe.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("$P pipelined:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"]["pipelined"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c22_bench_${P}_pipe.err").read()[-2500:])
PY
done
