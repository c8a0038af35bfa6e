#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/strict_gemm_probe.py --split16 --more > gpurun_out/r2_c17_strict_gemm.log 2>&1; cat gpurun_out/r2_c17_strict_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_c17_split16_gemm python tools/strict_gemm_probe.py --split16 > gpurun_out/r2_c17_ncu_strict.log 2>&1; tail -1 gpurun_out/r2_c17_ncu_strict.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c17_launches_fp32x3_fps4.csv \
  python tools/ncu_chain.py --fps 4 --step --precision fp32x3 > gpurun_out/r2_c17_ncu_list_strict.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_c17_bench_default.json 2> gpurun_out/r2_c17_bench_default.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c17_bench_default.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("default:", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3)); print(json.dumps(d["modes"]))
    print("parity", json.dumps({k: (v and {kk: v[kk] for kk in ("meets_bar", "logits_p99", "logits_max", "min_matched_frac")}) for k, v in d["parity"]["modes"].items()}))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c17_bench_default.err").read()[-1500:])
PY
cp gpurun_out/launch_times_fp32x3.json gpurun_out/r2_c17_launch_times_fp32x3.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -2
