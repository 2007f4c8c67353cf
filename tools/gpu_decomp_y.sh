#!/bin/bash
mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests/test_decomp.py -m gpu -q -x > gpurun_out/pytest_decomp.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decomp.log
  for g in 8 16; do timeout 600 python tools/bench_decomp.py --grid 1024 512 128 --tt 320 --mode y --ghost $g 2>&1 | tail -1; done | tee gpurun_out/decomp_y_n1.json
  timeout 600 python tools/bench_decomp.py --grid 1024 512 128 --tt 100 --mode x 2>&1 | tail -1 | tee -a gpurun_out/decomp_y_n1.json
else
  [ "$N" = "2" ] && { timeout 900 python -m pytest tests/test_decomp.py -m gpu -q -x -k two_gpu > gpurun_out/pytest_decomp_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decomp_n$N.log; }
  for g in ${GHOSTS:-16}; do timeout 900 python -m torch.distributed.run --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_decomp.py --grid ${GRID:-2048 1024 128} --tt ${TT:-320} --mode y --ghost $g 2>&1 | tail -1; done | tee gpurun_out/decomp_y_n$N.json
fi
