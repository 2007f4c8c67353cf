#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "field_fused or scatter_gradient" > gpurun_out/pytest_f12.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_f12.log
