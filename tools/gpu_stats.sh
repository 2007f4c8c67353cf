#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
ARGS="${LEAN_ARGS:---kernel systolic_lean}"
B200FDTD_LEAN_STATS=1 $B --tt 2000 --steps 1 --warmup 0 $ARGS > gpurun_out/lean_stats.log 2>&1
grep -c leanstats gpurun_out/lean_stats.log
fmt='import sys,json
j=json.loads(sys.stdin.read()); p=j["config"]["plan"]; print(round(j["value"],1),"Gcell/s", [round(20000*256*256*128/1e6/x,1) if False else x for x in []], p["tile_y"], p["stages"], j["clocks"])'
for i in 1 2 3 4 5; do $B --tt 4000 --steps 4 --warmup 1 $ARGS 2>&1 | tail -1 | python -c "$fmt"; done | tee gpurun_out/repeat_lean.log
for i in 1 2 3; do B200FDTD_LEAN_UNROLL=2 $B --tt 4000 --steps 4 --warmup 1 $ARGS 2>&1 | tail -1 | python -c "$fmt"; done | tee -a gpurun_out/repeat_lean.log
