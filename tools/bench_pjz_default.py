"""Throughput on pjz's DEFAULT engine geometry (use_reduced_precision=True, 128 - sum(pml) = 96
z-cells, /root/reference/src/pjz/_field.py:36-58) for every kernel that supports it."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pjz_b200 import fdtdz_jax
from tests.problems import random_problem

def run(kw, reps=2, **lp):
  kw = dict(kw); kw["launch_params"] = lp or None
  kw["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  try:
    fdtdz_jax.fdtdz(**kw); torch.cuda.synchronize()
  except Exception as e:
    return None, str(e)[:80]
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps): fdtdz_jax.fdtdz(**kw)
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / reps, fdtdz_jax.plan_info(**kw)

GEOMS = [(True, 96, (16, 16)), (False, 32, (16, 16)), (False, 48, (8, 8)), (True, 128, (16, 16)),
         (False, 64, (8, 8)), (True, 64, (16, 16))]
VARIANTS = [dict(), dict(kernel="systolic_lean"), dict(kernel="systolic_async", cols=1),
            dict(kernel="systolic_async", cols=2), dict(kernel="systolic"), dict(kernel="twopass")]
if os.environ.get("QUICK"):      # half-warp lean kernel against the AUTO plan, plus its tilings
  VARIANTS = [dict(), dict(kernel="systolic_async"), dict(kernel="systolic_lean")]
if os.environ.get("SWEEP"):      # protocol knobs of the sub-warp lean kernel (read at plan time)
  GEOMS = [(True, 96, (16, 16)), (True, 128, (16, 16)), (True, 64, (16, 16)), (False, 64, (8, 8))]
  VARIANTS = [dict(kernel="systolic_lean")]
  ENVS = [{}, {"B200FDTD_MAX_LEAD": "6"}, {"B200FDTD_MAX_LEAD": "16"}, {"B200FDTD_PF_AHEAD": "3"},
          {"B200FDTD_PF_AHEAD": "10"}, {"B200FDTD_SPIN_NS": "40"}, {"B200FDTD_SPIN_NS": "640"},
          {"B200FDTD_SVC_SLEEP": "50"}, {"B200FDTD_SVC_SLEEP": "800"}, {"B200FDTD_LEAN_DISCARD": "0"}]
else:
  ENVS = [{}]
for reduced, Z, pml in GEOMS:
  X = Y = 256
  tt = 2000
  kw = random_problem(domain=(X, Y, Z), sub=(X - 64, Y - 64, max(Z - 8, 1)), offset=(32, 32, 4), axis=0,
                      pml=pml, tt=tt, seed=1, output_steps=(tt - 1, tt, 1), reduced=reduced,
                      absorb_pad=32, absorb_coeff=1e-4)
  for lp, env in [(lp, env) for lp in VARIANTS for env in ENVS]:
    os.environ.update(env)
    ms, info = run(kw, **lp)
    for k in env: del os.environ[k]
    lp = dict(lp, **env)
    if ms is None:
      print(json.dumps({"reduced": reduced, "Z": Z, "lp": lp, "error": info})); continue
    print(json.dumps({"reduced": reduced, "grid": [X, Y, Z], "lp": lp, "gcell_s": X * Y * Z * tt / ms / 1e6,
                      "plan": {k: info[k] for k in ("kernel", "tile_y", "stages", "threads", "ctas")}}))
