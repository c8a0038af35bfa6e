#!/bin/bash
# First GPU call of round 2: everything that was written after the last GPU session of round 1 and is verified on the
# CPU only (host builds / op stand-ins). Order = cheapest evidence first; every step has its own timeout so a surprise in
# one does not eat the call. Usage:  gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
#                          and for the multi-GPU part:  gpurun --gpus 2 --timeout 900 -- 'WAVE=1 bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
if [ -z "$WAVE" ]; then
  # 1. new kernels + DFF + wavefront equivalence (tests/test_zz_train_ops_gpu.py), verbose so a partial run still tells
  timeout 900 python -m pytest tests/test_zz_train_ops_gpu.py -v --no-header -p no:cacheprovider > gpurun_out/r2_zz.log 2>&1
  echo "zz rc=$?" >> gpurun_out/r2_zz.log; grep -E "PASSED|FAILED|ERROR|rc=" gpurun_out/r2_zz.log | tail -30
  # 2. timing of the separable ROIAlign candidate and of the training-side kernels
  timeout 200 python tools/roi_probe.py > gpurun_out/r2_roi_probe.log 2>&1; tail -6 gpurun_out/r2_roi_probe.log
  # 2b. two key frames per launch of the per-frame branch against the default step
  for f in 1 2; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-strict --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_bench_fps$f.json 2> gpurun_out/r2_bench_fps$f.err
    python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_bench_fps$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("frames/step $f: no line", e)
PY
  done
  timeout 200 python tools/train_ops_probe.py > gpurun_out/r2_train_ops_probe.log 2>&1; tail -12 gpurun_out/r2_train_ops_probe.log
else
  # 3. wavefront schedule vs replicated-state schedule at N GPUs (N = number of visible devices)
  N=$(python -c "import torch; print(torch.cuda.device_count())")
  for mode in "--no-wave" ""; do        # "" = default: wavefront if every rank passes parallel.wave_selfcheck
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 30 --warmup 6 --skip-cpu-baseline --no-strict $mode > gpurun_out/r2_bench_${N}gpu${mode}.json 2> gpurun_out/r2_bench_${N}gpu${mode}.err
    echo "N=$N mode='$mode' rc=$?"; tail -2 gpurun_out/r2_bench_${N}gpu${mode}.err
    python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_bench_${N}gpu${mode}.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("N=${N} ${mode}", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("no line", e)
PY
  done
fi
du -sh gpurun_out
