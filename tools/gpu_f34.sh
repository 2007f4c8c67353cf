#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render.py tests/test_mode_gpu.py -m gpu -q -x > gpurun_out/pytest_f34.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_f34.log
F1_TT=2500 timeout 900 python tools/bench_postproc.py 2>&1 | tee gpurun_out/postproc.jsonl | tail -8
