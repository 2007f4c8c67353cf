#!/bin/bash
# Per-warp wait accounting of the sub-warp lean kernel (B200FDTD_LEAN_STATS=1) on cfg2 with fp16 storage.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
B200FDTD_LEAN_STATS=1 timeout 300 $B --tt 2000 --steps 1 --warmup 0 --reduced ${LEAN_ARGS} > gpurun_out/lean16_stats.log 2>&1
grep -c lean16stats gpurun_out/lean16_stats.log
grep "lean16stats" gpurun_out/lean16_stats.log | grep " t 0 " | sort -k3n -k7n | awk '{print}' | head -130
