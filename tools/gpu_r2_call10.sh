#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -q -p no:cacheprovider --no-header -rf -x -k "fp32x3 or split or stream_k or conv_matches or linear_matches" > gpurun_out/r2_c10_pytest_a.log 2>&1
echo "pytest A rc=$?"; tail -5 gpurun_out/r2_c10_pytest_a.log
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_full_gpu.py tests/test_zz_train_ops_gpu.py tests/test_extra_ops_gpu.py -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c10_pytest_b.log 2>&1
echo "pytest B rc=$?"; tail -5 gpurun_out/r2_c10_pytest_b.log
timeout 300 python tools/strict_probe.py 2 > gpurun_out/r2_c10_strict.log 2>&1; grep seg_len gpurun_out/r2_c10_strict.log
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline > gpurun_out/r2_c10_bench_strict.json 2> gpurun_out/r2_c10_bench_strict.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c10_bench_strict.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("strict fps4:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c10_bench_strict.err").read()[-1500:])
PY
cp gpurun_out/launch_times_fp32x3.json gpurun_out/r2_c10_launch_times_fp32x3.json 2>/dev/null
timeout 100 python tools/tf32_trunc_probe.py 2>&1 | tail -4
