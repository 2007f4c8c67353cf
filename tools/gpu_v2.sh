#!/bin/bash
mkdir -p gpurun_out
echo "== pytest async"; timeout 1200 python -m pytest tests -m gpu -q -k "async or large or plan" --maxfail=10 > gpurun_out/pytest_async.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_async.log
B="python bench.py --no-cpu --no-e2e"
echo "== sweep (tt=2000)"
for cfg in "--kernel systolic_async --prefetch 1" "--kernel systolic_async --prefetch 2" "--kernel systolic_async --prefetch 2 --stages 4" "--kernel systolic_async --prefetch 2 --stages 2" "--kernel systolic_async --prefetch 1 --stages 7" "--kernel systolic_async --prefetch 3 --tile-y 6" "--kernel systolic_async --prefetch 2 --tile-y 10"; do
  echo "-- $cfg"; timeout 200 $B --tt 2000 --steps 2 --warmup 1 $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), j['config']['plan'], j['clocks'])
except Exception as e: print('ERR',e)"
done | tee gpurun_out/sweep_v2.log
