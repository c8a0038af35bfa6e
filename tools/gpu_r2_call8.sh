#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c8_pytest.log; tail -6 gpurun_out/r2_c8_pytest.log
for f in 1 4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --precision f16 --no-parity --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_c8_bench_f$f.json 2> gpurun_out/r2_c8_bench_f$f.err
  cp gpurun_out/launch_times_f16.json gpurun_out/r2_c8_launch_times_f$f.json 2>/dev/null
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c8_bench_f$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "dom", d["roofline"]["dominant_kernel"]["name"], round(d["roofline"]["dominant_kernel"]["ms"],3))
except Exception as e:
    print("frames/step $f: no line", e); print(open("gpurun_out/r2_c8_bench_f$f.err").read()[-1500:])
PY
done
timeout 200 python tools/trace_backbone.py --fps 1 > gpurun_out/r2_c8_trace_fps1.log 2>&1; tail -19 gpurun_out/r2_c8_trace_fps1.log
timeout 200 python tools/trace_backbone.py --fps 1 --layer 56 > gpurun_out/r2_c8_trace_l56.log 2>&1; tail -22 gpurun_out/r2_c8_trace_l56.log
timeout 200 python tools/dcn_probe.py > gpurun_out/r2_c8_dcn_probe.log 2>&1; cat gpurun_out/r2_c8_dcn_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deform_im2col_tile|focal_loss" -s 6 -c 3 -o gpurun_out/r2_c8_dcn python tools/dcn_probe.py > gpurun_out/r2_c8_ncu_dcn.log 2>&1; tail -2 gpurun_out/r2_c8_ncu_dcn.log
for a in rdn_r101 fgfa_r101; do
  timeout 400 python bench.py --arch $a --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_c8_bench_$a.json 2> gpurun_out/r2_c8_bench_$a.err
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c8_bench_$a.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("$a:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$a: no line", e); print(open("gpurun_out/r2_c8_bench_$a.err").read()[-1500:])
PY
done
du -sh gpurun_out
