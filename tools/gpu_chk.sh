#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_decomp.py -m gpu -q -x -k "lean or large or long_run or y_session or projection" > gpurun_out/pytest_chk.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_chk.log
VARFILE=tools/variants.txt TT=6000 STEPS=2 tools/gpu_misc.sh
