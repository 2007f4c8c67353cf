#!/bin/bash
# One entry point for everything that is run on the GPU box through gpurun.  Each verb writes
# under gpurun_out/ (merged back into the container); copy what should be judged to profiles/.
#
#   tools/gpu.sh tests [pytest args]            pytest -m gpu (tail in gpurun_out/pytest_gpu.log)
#   tools/gpu.sh bench NAME [bench.py args]     one bench line -> gpurun_out/bench_NAME.json
#   tools/gpu.sh launches NAME [bench.py args]  ncu launch list (gpu__time_duration) of that command
#   tools/gpu.sh ncu NAME REGEX [bench.py args] ncu --set full + source of one launch of REGEX:
#                                               NAME.ncu-rep, NAME_raw.csv, NAME_source.csv, NAME_summary.json
#   tools/gpu.sh ab NAME VAR "v1 v2 .." [bench.py args]   A/B of an environment knob, one line each
#   tools/gpu.sh stats NAME [bench.py args]     B200FDTD_LEAN_STATS=1 per-warp wait accounting
#   tools/gpu.sh smoke                          __graft_entry__.smoke()
# Several verbs in one call:  tools/gpu.sh tests -- bench default -- ncu lean regex:lean_kernel --tt 400
mkdir -p gpurun_out
B="python bench.py"
short() { python -c "
import sys, json
for l in sys.stdin:
  l = l.strip()
  if not l.startswith('{'): continue
  j = json.loads(l); r = j.get('roofline') or {}
  print(round(j['value'], 2), j['unit'], 'frac', round(r.get('frac', 0), 3), 'e2e', (j.get('e2e') or {}).get('value'),
        j.get('plan') or j['config'].get('plan'), j.get('clocks'), {k: j[k] for k in ('reduced_precision', 'decomp') if k in j})"; }
run_one() {
  verb=$1; shift
  case $verb in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke.log;;
    bench) n=$1; shift; timeout 1200 $B "$@" > gpurun_out/bench_$n.log 2>&1; echo "bench $n rc=$?"; grep '^{' gpurun_out/bench_$n.log | tail -1 > gpurun_out/bench_$n.json; short < gpurun_out/bench_$n.json; grep -v '^{' gpurun_out/bench_$n.log | tail -3;;
    launches) n=$1; shift; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$n.csv $B "$@" > gpurun_out/launches_$n.log 2>&1; echo "launches $n rc=$?"; tail -4 gpurun_out/launches_$n.csv;;
    ncu) n=$1; k=$2; shift 2
      timeout 1500 ncu --set full --clock-control none --import-source on -k $k -c 1 -o gpurun_out/$n -f $B --no-cpu --no-e2e --steps 1 --warmup 0 "$@" > gpurun_out/ncu_$n.log 2>&1; echo "ncu $n rc=$?"; tail -2 gpurun_out/ncu_$n.log
      ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
      ncu -i gpurun_out/$n.ncu-rep --page source --csv > gpurun_out/${n}_source.csv 2>/dev/null
      python tools/ncu_summary.py gpurun_out/$n.ncu-rep ${CELL_UPDATES:-} > gpurun_out/${n}_summary.json 2>/dev/null
      ls -la gpurun_out/$n.ncu-rep;;
    ab) n=$1; var=$2; vals=$3; shift 3
      for v in $vals; do echo "-- $var=$v"; env $var=$v timeout 900 $B --no-cpu --no-e2e "$@" 2>&1 | grep '^{' | tail -1 | short; done | tee gpurun_out/ab_$n.log;;
    stats) n=$1; shift; B200FDTD_LEAN_STATS=1 timeout 900 $B --no-cpu --no-e2e --steps 1 --warmup 0 "$@" > gpurun_out/stats_$n.log 2>&1; echo "stats $n rc=$?"; grep -c stats gpurun_out/stats_$n.log;;
    *) echo "unknown verb $verb"; return 2;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" = "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader
