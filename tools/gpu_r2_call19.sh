#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_split16_gpu.py -m gpu -q -p no:cacheprovider --no-header -rf -k "tensor_memory" > gpurun_out/r2_c19_split16.log 2>&1
echo "a_tmem pytest rc=$?" >> gpurun_out/r2_c19_split16.log; tail -8 gpurun_out/r2_c19_split16.log
timeout 200 python tools/strict_gemm_probe.py --split16 --more > gpurun_out/r2_c19_gemm_ss.log 2>&1; cat gpurun_out/r2_c19_gemm_ss.log
timeout 200 python tools/strict_gemm_probe.py --split16 --more --a-tmem > gpurun_out/r2_c19_gemm_ts.log 2>&1; cat gpurun_out/r2_c19_gemm_ts.log
