#!/bin/bash
# ncu --set full (+source) of the half-warp lean kernel on the reduced-precision bend (256x256x128).
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
ARGS="${LEAN_ARGS:---kernel systolic_lean --reduced}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lean16 -c 1 -o gpurun_out/prof_lean16 -f $B --tt 200 --steps 1 --warmup 0 $ARGS > gpurun_out/ncu_full_lean16.log 2>&1; tail -3 gpurun_out/ncu_full_lean16.log
ncu -i gpurun_out/prof_lean16.ncu-rep --page source --csv > gpurun_out/prof_lean16_source.csv 2>/dev/null
ncu -i gpurun_out/prof_lean16.ncu-rep --page raw --csv > gpurun_out/prof_lean16_raw.csv 2>/dev/null
rm -f gpurun_out/prof_lean16.ncu-rep.tmp
$B --tt 4000 --steps 3 --warmup 1 $ARGS 2>&1 | tail -1 > gpurun_out/lean16_bench_line.json; cut -c1-600 gpurun_out/lean16_bench_line.json
