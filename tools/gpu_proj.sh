#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "projection" > gpurun_out/pytest_proj.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_proj.log
F1_TT=2500 python tools/bench_postproc.py 2>&1 | grep "f1 fused" | tee gpurun_out/f1.json
