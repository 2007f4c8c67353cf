#!/bin/bash
# Full round evidence: GPU test-suite, smoke, ncu launch list + full capture of the default plan,
# default bench line (full contract), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
B="python bench.py --no-cpu --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_lean.csv $B --tt 2000 --steps 2 --warmup 1 > gpurun_out/ncu_list_lean.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lean_kernel -c 1 -o gpurun_out/prof_lean -f $B --tt 400 --steps 1 --warmup 0 > gpurun_out/ncu_full_lean.log 2>&1; tail -2 gpurun_out/ncu_full_lean.log
ncu -i gpurun_out/prof_lean.ncu-rep --page source --csv > gpurun_out/prof_lean_source.csv 2>/dev/null
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log
F1_TT=2500 timeout 900 python tools/bench_postproc.py > gpurun_out/postproc.jsonl 2>&1; tail -3 gpurun_out/postproc.jsonl | cut -c1-300
