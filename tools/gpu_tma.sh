#!/bin/bash
# Parity + perf of the TMA-staged systolic kernel.
mkdir -p gpurun_out
echo "== smoke-size check"
timeout 120 python - <<'EOF' 2>&1 | tail -5
import numpy as np, torch
from oracle import fdtd_c
from pjz_b200 import fdtdz_jax
from tests.problems import random_problem
for dom, axis in [((24, 20, 32), 0), ((11, 23, 16), 1), ((10, 12, 128), 2)]:
    kw = random_problem(domain=dom, axis=axis, pml=(4, 6), tt=30, seed=11, output_steps=(20, 30, 4))
    want = fdtd_c.fdtdz(**kw)
    kw["launch_params"] = {"kernel": "systolic_tma"}
    kw["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
    out = fdtdz_jax.fdtdz(**kw).cpu().numpy()
    print(dom, axis, "bit-exact:", bool(np.array_equal(out, want)), "maxdiff", float(np.abs(out - want).max()))
EOF
echo "== pytest tma"; timeout 900 python -m pytest tests -m gpu -q -k "tma or large" --maxfail=5 > gpurun_out/pytest_tma.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_tma.log
B="python bench.py --no-cpu --no-e2e"
run() { echo "-- $ENVV $*"; timeout 300 env $ENVV $B --tt 2000 --steps 2 --warmup 1 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); p=j['config']['plan']; print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), p['kernel'],'tile',p['tile_y'],'stages',p['stages'],'thr',p['threads'],'ctas',p['ctas'],'smem',p['smem_bytes'],'D',p['prefetch'],'win',p['l2_window_mib'], 'W',j['clocks'].get('power_w_max'),'MHz',j['clocks'].get('sm_mhz'))
except Exception as e: print('ERR',e)"; }
{
ENVV=""
run --kernel systolic_tma
run --kernel systolic_tma --prefetch 2
run --kernel systolic_tma --stages 1
run --kernel systolic_tma --tile-y 10
ENVV="B200FDTD_PF_AHEAD=0"; run --kernel systolic_tma
ENVV="B200FDTD_MAX_LEAD=20"; run --kernel systolic_tma
ENVV=""; run --kernel systolic_async
} | tee gpurun_out/sweep_tma.log
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:systolic3_kernel -c 1 -o gpurun_out/prof_tma -f $B --tt 200 --steps 1 --warmup 0 --kernel systolic_tma > gpurun_out/ncu_full_tma.log 2>&1; tail -2 gpurun_out/ncu_full_tma.log
fi
