#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split16_gpu.py tests/test_engine_gpu.py tests/test_parity_full_gpu.py tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c18_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_c18_pytest.log | tail -8
for F in 4 8; do
timeout 400 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline --frames-per-step $F > gpurun_out/r2_c18_bench_strict_f$F.json 2> gpurun_out/r2_c18_bench_strict_f$F.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c18_bench_strict_f$F.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("strict fps$F:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c18_bench_strict_f$F.err").read()[-1500:])
PY
done
timeout 400 python bench.py --steps 10 --warmup 3 --precision f16 --no-parity --skip-cpu-baseline --frames-per-step 8 > gpurun_out/r2_c18_bench_f16_f8.json 2> gpurun_out/r2_c18_bench_f16_f8.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c18_bench_f16_f8.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("f16 fps8:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c18_bench_f16_f8.err").read()[-1500:])
PY
