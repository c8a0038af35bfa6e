#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split16_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c14_split16.log 2>&1
echo "split16 pytest rc=$?" >> gpurun_out/r2_c14_split16.log; tail -12 gpurun_out/r2_c14_split16.log
timeout 200 python tools/strict_gemm_probe.py --split16 > gpurun_out/r2_c14_strict_gemm.log 2>&1; tail -4 gpurun_out/r2_c14_strict_gemm.log
timeout 700 python -m pytest tests/test_parity_full_gpu.py tests/test_engine_gpu.py tests/test_extra_ops_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c14_parity.log 2>&1
echo "parity pytest rc=$?" >> gpurun_out/r2_c14_parity.log; tail -15 gpurun_out/r2_c14_parity.log
timeout 500 python tools/strict_probe.py 2 4 8 > gpurun_out/r2_c14_strict.log 2>&1; grep seg_len gpurun_out/r2_c14_strict.log || tail -20 gpurun_out/r2_c14_strict.log
MEGA_B200_SPLIT16_ATT=0 timeout 300 python tools/strict_probe.py 2 > gpurun_out/r2_c14_strict_noatt.log 2>&1; echo "without split16 attention:"; grep seg_len gpurun_out/r2_c14_strict_noatt.log || tail -20 gpurun_out/r2_c14_strict_noatt.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline > gpurun_out/r2_c14_bench_strict.json 2> gpurun_out/r2_c14_bench_strict.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c14_bench_strict.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("strict fps4:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c14_bench_strict.err").read()[-1500:])
PY
cp gpurun_out/launch_times_fp32x3.json gpurun_out/r2_c14_launch_times_fp32x3.json 2>/dev/null
