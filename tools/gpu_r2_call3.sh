#!/bin/bash
# Round 2, call 3: pin the tuned table, source-level profile of the chain kernel, new sub-module test
mkdir -p gpurun_out
timeout 1500 python tools/tune_all.py > gpurun_out/r2_c3_tune.log 2>&1; tail -12 gpurun_out/r2_c3_tune.log
cp gpurun_out/tuned_b200.json mega.pytorch_b200/mega_core/b200/tuned_b200.json
# launch list of one steady step (f16, graphs off so every kernel is a launch): shares, not absolutes
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. --csv --log-file gpurun_out/r2_c3_launches.csv \
  python bench.py --steps 2 --warmup 1 --precision f16 --no-parity --skip-cpu-baseline --skip-roofline --no-graph --prime 0 > gpurun_out/r2_c3_ncu_list.log 2>&1
tail -2 gpurun_out/r2_c3_ncu_list.log | cut -c1-300
# full capture with source of the first 95-layer chain launch (tuned table pinned -> no autotune launches before it)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -c 1 -o gpurun_out/r2_c3_chain \
  python bench.py --steps 1 --warmup 1 --precision f16 --no-parity --skip-cpu-baseline --skip-roofline --no-graph --prime 0 > gpurun_out/r2_c3_ncu_full.log 2>&1
tail -3 gpurun_out/r2_c3_ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
