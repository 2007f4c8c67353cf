#!/bin/bash
# ncu --set full (+source) of the lean kernel, then a repeatability check.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
ARGS="${LEAN_ARGS:---kernel systolic_lean --stages 7}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lean -c 1 -o gpurun_out/prof_lean -f $B --tt 200 --steps 1 --warmup 0 $ARGS > gpurun_out/ncu_full_lean.log 2>&1; tail -3 gpurun_out/ncu_full_lean.log
ncu -i gpurun_out/prof_lean.ncu-rep --page source --csv > gpurun_out/prof_lean_source.csv 2>/dev/null
ncu -i gpurun_out/prof_lean.ncu-rep --page raw --csv > gpurun_out/prof_lean_raw.csv 2>/dev/null
rm -f gpurun_out/prof_lean.ncu-rep.tmp
for i in 1; do B200FDTD_LEAN_STATS=1 $B --tt 2000 --steps 1 --warmup 0 $ARGS > gpurun_out/lean_stats.log 2>&1; $B --tt 4000 --steps 3 --warmup 1 $ARGS 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(round(j['value'],1),'Gcell/s', j['config']['plan'], j['clocks'])"; done | tee gpurun_out/repeat_lean.log
