#!/bin/bash
# First on-GPU check: smoke, parity tests, short benches.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest twopass"; timeout 900 python -m pytest tests -m gpu -q -k "twopass or host_path or xla or streams or workspace" --maxfail=20 > gpurun_out/pytest_twopass.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_twopass.log
echo "== pytest rest"; timeout 1200 python -m pytest tests -m gpu -q -k "not twopass" --maxfail=20 > gpurun_out/pytest_rest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_rest.log
echo "== bench twopass"; timeout 300 python bench.py --tt 2000 --steps 2 --warmup 1 --kernel twopass --no-cpu --no-e2e > gpurun_out/bench_twopass.log 2>&1; tail -2 gpurun_out/bench_twopass.log
echo "== bench systolic"; timeout 300 python bench.py --tt 2000 --steps 2 --warmup 1 --kernel systolic --no-cpu --no-e2e > gpurun_out/bench_systolic.log 2>&1; tail -2 gpurun_out/bench_systolic.log
