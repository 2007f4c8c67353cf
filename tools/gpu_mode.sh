#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mode_gpu.py -m gpu -q -x --durations=5 > gpurun_out/pytest_mode.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_mode.log
true
