#!/bin/bash
# round-validation on the GPU box: parity tests, bench lines, ncu launch list + full capture of a few dominant kernels.
# Everything written under gpurun_out/ must stay small (the whole directory is dropped above 64 MiB).
mkdir -p gpurun_out
PREC=${PREC:-f16}
timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 --precision $PREC > gpurun_out/bench_$PREC.json 2> gpurun_out/bench_$PREC.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_$PREC.json; tail -5 gpurun_out/bench_$PREC.err
cp gpurun_out/tuned_b200.json gpurun_out/tuned_b200_$PREC.json 2>/dev/null
if [ -n "$ALSO_TF32" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 --precision tf32 --no-strict --skip-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
  tail -c 1500 gpurun_out/bench_tf32.json
fi
export MEGA_B200_AUTOTUNE=0 MEGA_B200_TUNED=gpurun_out/tuned_b200_$PREC.json
# (1) per-launch durations of steady frames (no graph, so every kernel is its own launch)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches_$PREC.csv \
   python bench.py --steps 2 --warmup 1 --prime 2 --no-graph --no-strict --skip-cpu-baseline --precision $PREC > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
# (2) full-set captures, kept small: a window of conv_gemm launches of a steady frame, the soft-max, ROIAlign
BENCH_NCU="python bench.py --steps 1 --warmup 1 --prime 1 --no-graph --no-strict --skip-cpu-baseline --precision $PREC"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s ${NCU_SKIP:-1500} -c ${NCU_COUNT:-18} \
   -o gpurun_out/full_gemm_$PREC -f $BENCH_NCU > gpurun_out/ncu_full.log 2>&1
echo "ncu full gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relation_softmax|plain_softmax|roi_align" -s 70 -c 7 \
   -o gpurun_out/full_misc_$PREC -f $BENCH_NCU >> gpurun_out/ncu_full.log 2>&1
echo "ncu full misc rc=$?"
for f in gemm misc; do ncu -i gpurun_out/full_${f}_$PREC.ncu-rep --page raw --csv > gpurun_out/full_${f}_${PREC}_raw.csv 2>/dev/null; done
python - <<'PY'
import os, glob
for f in glob.glob("gpurun_out/*.ncu-rep"):
    if os.path.getsize(f) > 40 << 20:
        print("dropping", f, os.path.getsize(f)); os.remove(f)
PY
du -sh gpurun_out; ls -la gpurun_out
