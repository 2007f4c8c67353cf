#!/bin/bash
# Quick perf check of the default plan (+ variants given as extra "args" lines in $VARIANTS).
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $ENVV $*"; timeout 300 env $ENVV $B --tt ${TT:-2000} --steps 2 --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), p['kernel'],'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'],'D',p['prefetch'],'win',p['l2_window_mib'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
{
ENVV=""
run
while IFS= read -r line; do [ -n "$line" ] && run $line; done <<< "$VARIANTS"
} | tee gpurun_out/quick.log
if [ -n "$PYTEST" ]; then timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log; fi
