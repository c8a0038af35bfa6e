#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c26_launches_f16_fps4.csv \
  python tools/ncu_chain.py --fps 4 --step --precision f16 > gpurun_out/r2_c26_ncu_list_f16.log 2>&1
tail -2 gpurun_out/r2_c26_ncu_list_f16.log
