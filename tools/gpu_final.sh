#!/bin/bash
# Final check of a build: whole GPU suite, smoke, default bench line, reduced-precision sibling.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-260
timeout 200 python bench.py --reduced --no-cpu > gpurun_out/bench_reduced.log 2>&1; tail -1 gpurun_out/bench_reduced.log | cut -c1-260
