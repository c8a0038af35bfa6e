#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c11_pytest.log; tail -4 gpurun_out/r2_c11_pytest.log
timeout 200 python tools/strict_gemm_probe.py > gpurun_out/r2_c11_strict_gemm.log 2>&1; cat gpurun_out/r2_c11_strict_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_c11_strict_gemm python tools/strict_gemm_probe.py > gpurun_out/r2_c11_ncu_strict.log 2>&1; tail -1 gpurun_out/r2_c11_ncu_strict.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c11_launches_fp32x3_fps4.csv \
  python tools/ncu_chain.py --fps 4 --step --precision fp32x3 > gpurun_out/r2_c11_ncu_list_strict.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_c11_bench_default.json 2> gpurun_out/r2_c11_bench_default.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c11_bench_default.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("default:", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3)); print(json.dumps(d["modes"]))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c11_bench_default.err").read()[-1500:])
PY
ls -la gpurun_out/*.ncu-rep | tail -3
