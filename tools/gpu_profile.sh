#!/bin/bash
# ncu evidence + parameter sweep + one full default bench line.  Logs under gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e"
echo "== sweep (tt=1000)"
for cfg in "--kernel twopass" "--kernel systolic --stages 1" "--kernel systolic --stages 2" "--kernel systolic --stages 4" "--kernel systolic --stages 7" "--kernel systolic --tile-y 6 --stages 8" "--kernel systolic --tile-y 6 --stages 16" "--kernel systolic --threads 256" ; do
  echo "-- $cfg"; timeout 200 $B --tt 1000 --steps 2 --warmup 1 $cfg 2>&1 | tail -1 | python -c "
import sys,json
try:
  j=json.loads(sys.stdin.read()); print(round(j['value'],1),'Gcell/s frac',round(j['roofline']['frac'],3), j['config']['plan'], j['clocks'])
except Exception as e: print('ERR',e)"
done | tee gpurun_out/sweep.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_systolic.csv $B --tt 200 --steps 2 --warmup 1 --kernel systolic > gpurun_out/ncu_list_sys.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_twopass.csv $B --tt 200 --steps 2 --warmup 1 --kernel twopass > gpurun_out/ncu_list_two.log 2>&1
echo "== ncu full systolic"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:systolic_kernel -c 1 -o gpurun_out/prof_systolic -f $B --tt 200 --steps 1 --warmup 0 --kernel systolic > gpurun_out/ncu_full_sys.log 2>&1; tail -3 gpurun_out/ncu_full_sys.log
echo "== ncu full twopass"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:twopass -s 40 -c 2 -o gpurun_out/prof_twopass -f $B --tt 100 --steps 1 --warmup 0 --kernel twopass > gpurun_out/ncu_full_two.log 2>&1; tail -3 gpurun_out/ncu_full_two.log
echo "== default bench (full contract)"
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log
ls -la gpurun_out
