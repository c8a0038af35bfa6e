#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split16_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c15_split16.log 2>&1
rc=$?; echo "split16 pytest rc=$rc" >> gpurun_out/r2_c15_split16.log; tail -12 gpurun_out/r2_c15_split16.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 200 python tools/strict_gemm_probe.py --split16 > gpurun_out/r2_c15_strict_gemm.log 2>&1; tail -4 gpurun_out/r2_c15_strict_gemm.log
timeout 700 python -m pytest tests/test_parity_full_gpu.py tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c15_parity.log 2>&1
echo "parity pytest rc=$?" >> gpurun_out/r2_c15_parity.log; tail -8 gpurun_out/r2_c15_parity.log
timeout 500 python tools/strict_probe.py 2 4 > gpurun_out/r2_c15_strict.log 2>&1; grep seg_len gpurun_out/r2_c15_strict.log || tail -20 gpurun_out/r2_c15_strict.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline > gpurun_out/r2_c15_bench_strict.json 2> gpurun_out/r2_c15_bench_strict.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c15_bench_strict.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("strict fps4:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c15_bench_strict.err").read()[-1500:])
PY
cp gpurun_out/launch_times_fp32x3.json gpurun_out/r2_c15_launch_times_fp32x3.json 2>/dev/null
