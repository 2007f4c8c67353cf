#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_cabi.py -m gpu -q -x -k "host or projection or xla or workspace or golden" > gpurun_out/pytest_e2e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_e2e.log
python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('value',round(j['value'],2),'e2e',round(j['e2e']['value'],2), j['clocks'])"
