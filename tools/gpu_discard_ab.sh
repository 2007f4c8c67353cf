#!/bin/bash
# A/B of the L2 discard on full-length runs (20000 steps): flagship fp32 kernel and the sub-warp kernel.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3"
show() { tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(round(j['value'],2),'Gcell/s frac',round(j['roofline']['frac'],3), j['config']['plan']['kernel'], j['clocks'])"; }
{
for d in 1 0 1 0; do echo "fp32 cfg2 discard=$d"; B200FDTD_LEAN_DISCARD=$d timeout 200 $B 2>&1 | show; done
for d in 0 1 0 1; do echo "fp16 cfg2 discard=$d"; B200FDTD_LEAN_DISCARD=$d timeout 200 $B --reduced 2>&1 | show; done
} | tee gpurun_out/discard_ab.log
