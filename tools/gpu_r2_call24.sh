#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_train_ops_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -k "pipelined" > gpurun_out/r2_c24_pipe.log 2>&1
echo "pipelined pytest rc=$?" >> gpurun_out/r2_c24_pipe.log; tail -5 gpurun_out/r2_c24_pipe.log
timeout 900 python bench.py > gpurun_out/r2_c24_bench_default.json 2> gpurun_out/r2_c24_bench_default.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c24_bench_default.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("default:", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), d["config"]["pipelined"], d["e2e"]["api"]); print(json.dumps(d["modes"]))
    print("parity", json.dumps({k: (v and {kk: v[kk] for kk in ("meets_bar", "logits_p99", "logits_max", "min_matched_frac")}) for k, v in d["parity"]["modes"].items()}))
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c24_bench_default.err").read()[-1500:])
PY
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32x3 --no-parity --skip-cpu-baseline --skip-roofline --pipeline 0 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; d=json.loads(l); print('strict, not pipelined:', round(d['value'],1), round(d['e2e']['value'],1))"
