#!/bin/bash
# A/B of the split cp.async issue (B200FDTD_LEAN_SPLIT) on the fp32 flagship and the sub-warp kernel.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 2 --tt ${TT:-6000}"
show() { tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(round(j['value'],2),'Gcell/s frac',round(j['roofline']['frac'],3), j['config']['plan']['kernel'], j['clocks']['sm_mhz'], j['clocks'].get('power_w_max'))"; }
{
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean" 2>&1 | tail -2
B200FDTD_LEAN_SPLIT=1 timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean" 2>&1 | tail -2
for d in 0 1 0 1; do echo "fp32 cfg2 split=$d"; B200FDTD_LEAN_SPLIT=$d timeout 200 $B 2>&1 | show; done
for d in 0 1 0 1; do echo "fp16 cfg2 split=$d"; B200FDTD_LEAN_SPLIT=$d timeout 200 $B --reduced 2>&1 | show; done
} | tee gpurun_out/split_ab.log
