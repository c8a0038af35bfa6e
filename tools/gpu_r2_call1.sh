#!/bin/bash
# Round 2, first GPU call: suite status + the measurements the round's decisions hang on.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
# 1. whole GPU suite (no -x here: this call wants every verdict)
timeout 1000 python -m pytest tests -m gpu -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c1_pytest.log; tail -15 gpurun_out/r2_c1_pytest.log
cp gpurun_out/parity_full.json gpurun_out/r2_c1_parity_full.json 2>/dev/null
cp gpurun_out/engine_parity.json gpurun_out/r2_c1_engine_parity.json 2>/dev/null
# 2. key frames per launch of the per-frame branch
for f in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-strict --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_c1_bench_fps$f.json 2> gpurun_out/r2_c1_bench_fps$f.err
  cp gpurun_out/launch_times.json gpurun_out/r2_c1_launch_times_fps$f.json 2>/dev/null
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c1_bench_fps$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("frames/step $f: no line", e)
PY
done
# 3. strict mode: per-launch breakdown
timeout 300 python bench.py --steps 10 --warmup 3 --no-strict --skip-cpu-baseline --precision fp32x3 > gpurun_out/r2_c1_bench_fp32x3.json 2> gpurun_out/r2_c1_bench_fp32x3.err
cp gpurun_out/launch_times.json gpurun_out/r2_c1_launch_times_fp32x3.json 2>/dev/null
tail -c 600 gpurun_out/r2_c1_bench_fp32x3.json
# 4. in-kernel timeline of the chain kernel
timeout 120 python tools/trace_chain.py --cta 0 --blocks 4 > gpurun_out/r2_c1_trace.log 2>&1; head -3 gpurun_out/r2_c1_trace.log
# 5. ROIAlign candidates
timeout 100 python tools/roi_probe.py > gpurun_out/r2_c1_roi.log 2>&1
MEGA_B200_ROI_SEPARABLE=1 timeout 100 python tools/roi_probe.py >> gpurun_out/r2_c1_roi.log 2>&1; cat gpurun_out/r2_c1_roi.log
# 6. where the strict mode loses its digits
timeout 400 python tools/diag_parity.py > gpurun_out/r2_c1_diag.log 2>&1; grep -E "AUDIT|fp32x3" gpurun_out/r2_c1_diag.log | head -20
du -sh gpurun_out
