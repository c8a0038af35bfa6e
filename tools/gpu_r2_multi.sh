#!/bin/bash
# multi-GPU bench through the driver's own launch line; N = visible GPUs
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/r2_multi_${N}gpu.json 2> gpurun_out/r2_multi_${N}gpu.err
echo "N=$N rc=$?"; tail -3 gpurun_out/r2_multi_${N}gpu.err
python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_multi_${N}gpu.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l)
    print("N=${N}", "head", d["config"]["precision"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["config"]["schedule"], d["config"]["wave_selfcheck"])
    print("modes", json.dumps(d["modes"]))
    print("per_rank", json.dumps(d.get("per_rank")))
    print("comm", json.dumps(d.get("comm")))
    print("clocks", json.dumps(d.get("clocks")))
    print("parity", json.dumps({k: (v and {kk: v[kk] for kk in ("meets_bar", "logits_p99", "logits_max", "min_matched_frac")}) for k, v in d["parity"]["modes"].items()}))
except Exception as e:
    print("no line", e)
PY
