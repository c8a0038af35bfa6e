#!/bin/bash
# scaling check on one multi-GPU box: bench.py at N = 1, 2, 4 ... (N list in $NS), same flags as the driver's launch
mkdir -p gpurun_out
for n in ${NS:-1 2 4}; do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --skip-cpu-baseline --no-strict > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu-baseline --no-strict > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  fi
  echo "N=$n rc=$?"; tail -3 gpurun_out/bench_${n}gpu.err
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/bench_${n}gpu.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("N=${n}", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("no line", e)
PY
done
cp gpurun_out/tuned_b200.json gpurun_out/tuned_b200_scale.json 2>/dev/null
