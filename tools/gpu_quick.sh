#!/bin/bash
# quick GPU check: parity tests + one bench line (no profiler)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
if [ -z "$NO_BENCH" ]; then
timeout 900 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -c 2600 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
fi
