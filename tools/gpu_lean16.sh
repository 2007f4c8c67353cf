#!/bin/bash
# Sub-warp lean kernel: parity tests first, then throughput on pjz's default geometries.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "halfwarp or subwarp or rejects" --maxfail=4 > gpurun_out/lean16_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/lean16_tests.log
if [ -n "$SWEEP" ]; then
  timeout 300 python tools/bench_pjz_default.py > gpurun_out/lean16_sweep.jsonl 2> gpurun_out/lean16_sweep.err
  echo "sweep rc=$?"; cut -c1-200 gpurun_out/lean16_sweep.jsonl; tail -3 gpurun_out/lean16_sweep.err
else
  QUICK=1 timeout 300 python tools/bench_pjz_default.py > gpurun_out/lean16_bench.jsonl 2> gpurun_out/lean16_bench.err
  echo "bench rc=$?"; cut -c1-330 gpurun_out/lean16_bench.jsonl; tail -3 gpurun_out/lean16_bench.err
fi
