#!/bin/bash
# Round 2, call 2: interleaved (depth-2) chains, key frames per step, strict-mode segment length, new bench.py
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_ops_gpu.py "tests/test_zz_train_ops_gpu.py::test_two_key_frames_per_call_on_device" "tests/test_engine_gpu.py::test_mega_r101_f16_matches_reference_fixture" -q -p no:cacheprovider --no-header -rf -x > gpurun_out/r2_c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c2_pytest.log; tail -8 gpurun_out/r2_c2_pytest.log
for dual in 1 0; do for f in 1 2 4; do
  MEGA_B200_DUAL_CHAIN=$dual timeout 300 python bench.py --steps 20 --warmup 5 --precision f16 --no-parity --skip-cpu-baseline --frames-per-step $f > gpurun_out/r2_c2_bench_d${dual}_f$f.json 2> gpurun_out/r2_c2_bench_d${dual}_f$f.err
  cp gpurun_out/launch_times_f16.json gpurun_out/r2_c2_launch_times_d${dual}_f$f.json 2>/dev/null
  python - <<PY
import json
try:
    l = [x for x in open("gpurun_out/r2_c2_bench_d${dual}_f$f.json").read().splitlines() if x.startswith("{")][-1]
    d = json.loads(l); print("dual $dual frames/step $f:", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roof", round(d["roofline"]["frac"], 3), "dom", d["roofline"]["dominant_kernel"])
except Exception as e:
    print("dual $dual frames/step $f: no line", e); print(open("gpurun_out/r2_c2_bench_d${dual}_f$f.err").read()[-1500:])
PY
done; done
timeout 500 python tools/strict_probe.py 4 2 1 > gpurun_out/r2_c2_strict.log 2>&1; grep seg_len gpurun_out/r2_c2_strict.log
du -sh gpurun_out
