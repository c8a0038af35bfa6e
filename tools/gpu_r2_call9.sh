#!/bin/bash
# evidence call: tests on the final kernels, smoke(), benches, ncu captures (launch list + full captures with source)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c9_pytest.log; tail -4 gpurun_out/r2_c9_pytest.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for f in 1 4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --precision f16 --no-parity --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_c9_bench_f$f.json 2> gpurun_out/r2_c9_bench_f$f.err
  cp gpurun_out/launch_times_f16.json gpurun_out/r2_c9_launch_times_f$f.json 2>/dev/null
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c9_bench_f$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "dom", d["roofline"]["dominant_kernel"]["name"], round(d["roofline"]["dominant_kernel"]["ms"],3))
except Exception as e:
    print("frames/step $f: no line", e); print(open("gpurun_out/r2_c9_bench_f$f.err").read()[-1500:])
PY
done
timeout 200 python tools/trace_backbone.py --fps 4 > gpurun_out/r2_c9_trace_fps4.log 2>&1; tail -20 gpurun_out/r2_c9_trace_fps4.log
timeout 200 python tools/dcn_probe.py > gpurun_out/r2_c9_dcn_probe.log 2>&1; cat gpurun_out/r2_c9_dcn_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deform_im2col_tile|focal_loss" -s 6 -c 3 -o gpurun_out/r2_c9_dcn python tools/dcn_probe.py > gpurun_out/r2_c9_ncu_dcn.log 2>&1; tail -1 gpurun_out/r2_c9_ncu_dcn.log
for a in fgfa_r101; do
  timeout 400 python bench.py --arch $a --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r2_c9_bench_$a.json 2> gpurun_out/r2_c9_bench_$a.err
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c9_bench_$a.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("$a:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$a: no line", e); print(open("gpurun_out/r2_c9_bench_$a.err").read()[-1500:])
PY
done
# launch list of one steady step at 4 key frames per step (profiler on for that step only): shares, not absolutes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c9_launches_f16_fps4.csv \
  python tools/ncu_chain.py --fps 4 --step > gpurun_out/r2_c9_ncu_list.log 2>&1
tail -1 gpurun_out/r2_c9_ncu_list.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c9_launches_fp32x3_fps4.csv \
  python tools/ncu_chain.py --fps 4 --step --precision fp32x3 > gpurun_out/r2_c9_ncu_list_strict.log 2>&1
# full capture (source-level) of the 190-layer backbone chain of a step
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_c9_chain_fps4 \
  python tools/ncu_chain.py --fps 4 > gpurun_out/r2_c9_ncu_full.log 2>&1
tail -2 gpurun_out/r2_c9_ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
