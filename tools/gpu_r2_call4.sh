#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_backbone.py --fps 1 > gpurun_out/r2_c4_trace_fps1.log 2>&1; cat gpurun_out/r2_c4_trace_fps1.log | tail -32
timeout 300 python tools/trace_backbone.py --fps 4 > gpurun_out/r2_c4_trace_fps4.log 2>&1; cat gpurun_out/r2_c4_trace_fps4.log | tail -32
timeout 300 python -m pytest tests/test_extra_ops_gpu.py -q -p no:cacheprovider --no-header -x 2>&1 | tail -5
