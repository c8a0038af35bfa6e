#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_split16_gpu.py tests/test_parity_full_gpu.py -m gpu -x -q -p no:cacheprovider --no-header -rf > gpurun_out/r2_c27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_c27_pytest.log; grep -n "FAILED\|passed\|failed\|rc=" gpurun_out/r2_c27_pytest.log | tail -5
for P in f16 fp32x3; do
timeout 200 python bench.py --steps 10 --warmup 3 --precision $P --no-parity --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "
import json,sys
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]; d=json.loads(l); print('$P:', round(d['value'],1), round(d['e2e']['value'],1))"
done
