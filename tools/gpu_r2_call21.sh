#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c21_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c21_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_c21_pytest.log | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_c21_smoke.log 2>&1; tail -3 gpurun_out/r2_c21_smoke.log
for i in 1 2 3 4 5; do
timeout 300 python -m pytest tests/test_parity_full_gpu.py tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider --no-header -k "fp32x3 or do_not_depend_on_world_size" 2>&1 | tail -1
done
timeout 900 python bench.py > gpurun_out/r2_c21_bench_default.json 2> gpurun_out/r2_c21_bench_default.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c21_bench_default.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("default:", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"]); print(json.dumps(d["modes"]))
    print("parity", json.dumps({k: (v and {kk: v[kk] for kk in ("meets_bar", "logits_p99", "logits_max", "min_matched_frac")}) for k, v in d["parity"]["modes"].items()}))
    print("cpu", json.dumps(d.get("cpu_baseline")), "launches", d.get("gpu_launches"), "clocks", json.dumps(d.get("clocks")))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c21_bench_default.err").read()[-1500:])
PY
