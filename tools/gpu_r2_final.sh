#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_final_pytest.log | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_final_bench.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("default:", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), d["config"]["pipelined"]); print(json.dumps(d["modes"]))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_final_bench.err").read()[-1500:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; tail -c 600 gpurun_out/r2_final_ref.json
