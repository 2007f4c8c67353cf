#!/bin/bash
# Lean-kernel check: parity tests first, then a small sweep on cfg2 (extra variants in $VARIANTS).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean or large" > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_lean.log
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $*"; envs=(); while [[ "$1" == *=* ]]; do envs+=("$1"); shift; done; timeout 300 env "${envs[@]}" $B --tt ${TT:-2000} --steps 2 --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), p['kernel'],'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'],'D',p['prefetch'],'win',p['l2_window_mib'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
{
ENVV=""
run --kernel systolic_lean --tt 1000
run --kernel systolic_lean
while IFS= read -r line; do [ -n "$line" ] && run $line; done < ${VARFILE:-/dev/null}
} | tee gpurun_out/quick_lean.log
