#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "async or golden or stress or reduced or large" > gpurun_out/pytest_async.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_async.log
python tools/bench_pjz_default.py 2>&1 | grep -v '"kernel": "twopass"\|"lp": {"kernel": "systolic"}' | cut -c1-200 | tee gpurun_out/pjz_default2.jsonl
printf -- "--kernel systolic_async\n--kernel systolic_async --reduced\n" > /tmp/v2.txt
VARFILE=/tmp/v2.txt TT=4000 STEPS=2 tools/gpu_misc.sh
