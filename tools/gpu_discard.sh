#!/bin/bash
mkdir -p gpurun_out
B200FDTD_LEAN_DISCARD=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean or large" > gpurun_out/pytest_discard.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_discard.log
B="python bench.py --no-cpu --no-e2e"
fmt='import sys,json
j=json.loads(sys.stdin.read()); p=j["config"]["plan"]; print(round(j["value"],1),"Gcell/s", p["tile_y"], p["stages"], j["clocks"])'
for d in 0 1 0 1; do echo "discard=$d"; B200FDTD_LEAN_DISCARD=$d $B --tt 4000 --steps 3 --warmup 1 --kernel systolic_lean 2>&1 | tail -1 | python -c "$fmt"; done | tee gpurun_out/discard.log
echo demux; for d in 0 1; do B200FDTD_LEAN_DISCARD=$d $B --tt 2000 --steps 2 --warmup 1 --kernel systolic_lean --workload demux 2>&1 | tail -1 | python -c "$fmt"; done | tee -a gpurun_out/discard.log
B200FDTD_LEAN_DISCARD=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:lean_kernel -c 1 $B --tt 400 --steps 1 --warmup 0 --kernel systolic_lean 2>&1 | grep -E "dram__|gpu__time|hit_rate" | tee -a gpurun_out/discard.log
