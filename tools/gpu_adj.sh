#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render.py tests/test_mode_gpu.py -m gpu -q -x > gpurun_out/pytest_f34.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_f34.log
timeout 900 python tools/bench_adjoint_step.py --tt 20000 2>&1 | tail -2 | tee gpurun_out/adjoint_step.json
python tools/bench_postproc.py 2>&1 | grep "f4 renderer" | cut -c1-400
