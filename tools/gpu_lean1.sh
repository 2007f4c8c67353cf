#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "one_column or large" > gpurun_out/pytest_lean1.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_lean1.log
VARFILE=tools/variants.txt TT=4000 tools/gpu_misc.sh
