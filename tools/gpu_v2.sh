#!/bin/bash
# Parity tests of the async kernel + a small tuning sweep (+ optional ncu capture: NCU=1).
mkdir -p gpurun_out
echo "== pytest async"; timeout 1200 python -m pytest tests -m gpu -q -k "async or large or plan" --maxfail=10 > gpurun_out/pytest_async.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_async.log
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $ENVV $*"; timeout 200 env $ENVV $B --tt 2000 --steps 2 --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), 'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'],'D',p['prefetch'],'win',p['l2_window_mib'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
echo "== sweep (tt=2000)"
{
ENVV=""
run --kernel systolic_async --prefetch 1
run --kernel systolic_async --prefetch 2
run --kernel systolic_async --prefetch 1 --stages 1
run --kernel systolic_async --prefetch 1 --tile-y 6
run --kernel systolic_async --prefetch 1 --tile-y 4
ENVV="B200FDTD_SVC_SLEEP=1000"; run --kernel systolic_async --prefetch 1
ENVV="B200FDTD_SVC_SLEEP=100"; run --kernel systolic_async --prefetch 1
ENVV="B200FDTD_MAX_LEAD=20"; run --kernel systolic_async --prefetch 1
ENVV="B200FDTD_MAX_LEAD=20 B200FDTD_SVC_SLEEP=1000"; run --kernel systolic_async --prefetch 1 --tile-y 6
} | tee gpurun_out/sweep_v6.log
if [ -n "$NCU" ]; then
echo "== ncu full async"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:systolic2_kernel -c 1 -o gpurun_out/prof_async -f $B --tt 200 --steps 1 --warmup 0 --kernel systolic_async --prefetch 1 > gpurun_out/ncu_full_async.log 2>&1; tail -2 gpurun_out/ncu_full_async.log
fi
