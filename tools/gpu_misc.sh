#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $*"; envs=(); while [[ "$1" == *=* ]]; do envs+=("$1"); shift; done; timeout 300 env "${envs[@]}" $B --tt ${TT:-2000} --steps ${STEPS:-2} --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), j['config']['grid'], j['dtype'], p['kernel'],'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
while IFS= read -r line; do [ -n "$line" ] && run $line; done < ${VARFILE:-/dev/null} | tee gpurun_out/misc.log
