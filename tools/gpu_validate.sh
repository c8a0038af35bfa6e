#!/bin/bash
# round-validation on the GPU box: parity tests, bench lines (b200 + reference arm), ncu launch list + a full capture of
# the dominant kernels. Everything written under gpurun_out/ must stay small (the directory is dropped above 64 MiB).
mkdir -p gpurun_out
PREC=${PREC:-f16}
timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --precision $PREC > gpurun_out/bench_$PREC.json 2> gpurun_out/bench_$PREC.err; echo "bench rc=$?"
tail -c 3800 gpurun_out/bench_$PREC.json; tail -3 gpurun_out/bench_$PREC.err
cp gpurun_out/tuned_b200.json gpurun_out/tuned_b200_$PREC.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 600 gpurun_out/bench_reference.json
export MEGA_B200_AUTOTUNE=0 MEGA_B200_TUNED=gpurun_out/tuned_b200_$PREC.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 160 --csv --log-file gpurun_out/launches_$PREC.csv \
   python bench.py --steps 2 --warmup 1 --prime 2 --no-graph --no-strict --skip-cpu-baseline --precision $PREC > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_chain|roi_align|relation_softmax|conv_gemm_kernel|greedy" -s 40 -c 12 \
   -o gpurun_out/full_$PREC -f python bench.py --steps 1 --warmup 1 --prime 1 --no-graph --no-strict --skip-cpu-baseline --precision $PREC > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/full_$PREC.ncu-rep --page raw --csv > gpurun_out/full_${PREC}_raw.csv 2>/dev/null
python - <<'PY'
import os, glob
for f in glob.glob("gpurun_out/*.ncu-rep"):
    if os.path.getsize(f) > 30 << 20:
        print("dropping", f, os.path.getsize(f)); os.remove(f)
PY
du -sh gpurun_out
