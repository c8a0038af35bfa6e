#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider --no-header -k "do_not_depend_on_world_size" 2>&1 | tail -3
done
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c16_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_c16_pytest.log | tail -12
