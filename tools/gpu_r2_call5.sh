#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_engine_gpu.py tests/test_parity_full_gpu.py -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c5_pytest.log; tail -6 gpurun_out/r2_c5_pytest.log
for dual in 0 1; do for f in 1 2 4; do
  MEGA_B200_DUAL_CHAIN=$dual timeout 300 python bench.py --steps 20 --warmup 5 --precision f16 --no-parity --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_c5_bench_d${dual}_f$f.json 2> gpurun_out/r2_c5_bench_d${dual}_f$f.err
  cp gpurun_out/launch_times_f16.json gpurun_out/r2_c5_launch_times_d${dual}_f$f.json 2>/dev/null
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c5_bench_d${dual}_f$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("dual $dual frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "dom", d["roofline"]["dominant_kernel"]["name"], round(d["roofline"]["dominant_kernel"]["ms"],3))
except Exception as e:
    print("dual $dual frames/step $f: no line", e); print(open("gpurun_out/r2_c5_bench_d${dual}_f$f.err").read()[-1500:])
PY
done; done
timeout 200 python tools/trace_backbone.py --fps 1 > gpurun_out/r2_c5_trace_fps1.log 2>&1; tail -20 gpurun_out/r2_c5_trace_fps1.log
