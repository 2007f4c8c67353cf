#!/bin/bash
# Whole GPU suite + smoke + the default bench line and its reduced-precision sibling.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
QUICK=1 timeout 300 python tools/bench_pjz_default.py > gpurun_out/lean16_bench.jsonl 2> gpurun_out/lean16_bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/lean16_bench.jsonl; tail -3 gpurun_out/lean16_bench.err
if [ -z "$NOBENCH" ]; then
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-900
timeout 300 python bench.py --reduced --no-cpu --tt 8000 > gpurun_out/bench_reduced.log 2>&1; tail -1 gpurun_out/bench_reduced.log | cut -c1-900
fi
