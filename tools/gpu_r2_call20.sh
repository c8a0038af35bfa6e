#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/tune_all.py > gpurun_out/r2_c20_tune.log 2>&1; tail -8 gpurun_out/r2_c20_tune.log
ls -la gpurun_out/tuned_b200.json
