#!/bin/bash
# Full GPU parity tests + a small tuning sweep (+ optional ncu capture: NCU=1).
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $ENVV $*"; timeout 200 env $ENVV $B --tt 2000 --steps 2 --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), 'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'],'D',p['prefetch'],'win',p['l2_window_mib'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
echo "== sweep (tt=2000)"
{
ENVV=""
run --kernel twopass
run --kernel systolic
run --kernel systolic_async --prefetch 1
run --kernel systolic_async --prefetch 1 --stages 1
run --kernel systolic_async --prefetch 1 --stages 5
run --kernel systolic_async --prefetch 1 --tile-y 10
run --kernel systolic_async --prefetch 1 --tile-y 8
run --kernel systolic_async --prefetch 2
ENVV="B200FDTD_PF_AHEAD=0"; run --kernel systolic_async --prefetch 1
ENVV="B200FDTD_MAX_LEAD=16"; run --kernel systolic_async --prefetch 1
} | tee gpurun_out/sweep_v7.log
if [ -n "$NCU" ]; then
echo "== ncu full async"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:systolic2_kernel -c 1 -o gpurun_out/prof_async -f $B --tt 200 --steps 1 --warmup 0 --kernel systolic_async --prefetch 1 > gpurun_out/ncu_full_async.log 2>&1; tail -2 gpurun_out/ncu_full_async.log
fi
