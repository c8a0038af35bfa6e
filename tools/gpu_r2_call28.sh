#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -x -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c28_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c28_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_c28_pytest.log | tail -5
