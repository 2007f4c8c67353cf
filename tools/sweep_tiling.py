"""Tile width x stage count sweep of the (sub-)warp kernels on the 256 x 256 bend (GPU).

For every case the default plan runs first; every explicit (tile_y, stages) plan is then timed on
the same inputs AND its snapshots are compared bit for bit with the default plan's (all plans are
bit-exact against the oracle, so they must agree with each other).  One JSON line per plan.

  python tools/sweep_tiling.py [--tt 2000] [--out gpurun_out/sweep_tiling.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pjz_b200 import _field as glue                      # noqa: E402
from pjz_b200 import fdtdz_jax                           # noqa: E402
from pjz_b200 import workloads as W                      # noqa: E402

CASES = [
    # label, z cells, reduced, [(tile_y, stages) ...]   (0 = the planner's choice)
    ("fp16 z96", 96, True, [(19, 0), (17, 0), (15, 0), (13, 0), (11, 0), (9, 0), (7, 0), (5, 0),
                            (0, 8), (0, 10), (15, 6), (7, 8)]),
    ("fp16 z128", 128, True, [(19, 0), (15, 0), (11, 0), (7, 0), (0, 8)]),
    ("fp32 z128", 128, False, [(13, 0), (11, 0), (9, 0), (7, 0), (0, 6)]),
]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--tt", type=int, default=2000)
  ap.add_argument("--out", default="gpurun_out/sweep_tiling.jsonl")
  ap.add_argument("--budget", type=float, default=60.0, help="stop starting new plans after this many seconds")
  args = ap.parse_args()
  os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
  t_start = time.time()
  with open(args.out, "w") as fh:
    for label, zz, reduced, plans in CASES:
      eps, ports, params, omega = W.bend(total=(256, 256, zz), reduced=reduced)
      params = params._replace(tt=args.tt)
      axis, pos, _ = ports[0]
      kw, _, _ = glue.engine_inputs(eps, W.gaussian_port_source(eps, axis, pos), omega, pos, params)
      kw = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in kw.items()}
      warm = dict(kw, source_waveform=kw["source_waveform"][:200].contiguous()
                  if hasattr(kw["source_waveform"], "contiguous") else kw["source_waveform"][:200],
                  output_steps=(199, 200, 1))
      base = None
      for tile_y, stages in [(0, 0)] + plans:
        if time.time() - t_start > args.budget:
          break
        lp = {"kernel": "systolic_lean"}
        if tile_y:
          lp["tile_y"] = tile_y
        if stages:
          lp["stages"] = stages
        rec = {"case": label, "tile_y_req": tile_y, "stages_req": stages}
        try:
          k2 = dict(kw, launch_params=lp)
          rec["plan"] = fdtdz_jax.plan_info(**k2)
          fdtdz_jax.fdtdz(**dict(warm, launch_params=lp))
          torch.cuda.synchronize()
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record()
          out = fdtdz_jax.fdtdz(**k2)
          b.record()
          torch.cuda.synchronize()
          rec["gcell_s"] = 256 * 256 * zz * args.tt / (a.elapsed_time(b) / 1e3) / 1e9
          rec["finite"] = bool(torch.isfinite(out).all())
          if base is None:
            base = out
            rec["same_bits_as_default"] = True
          else:
            rec["same_bits_as_default"] = bool(torch.equal(out, base))
        except Exception as e:                               # noqa: BLE001
          rec["error"] = f"{type(e).__name__}: {str(e)[:160]}"
        fh.write(json.dumps(rec) + "\n")
        fh.flush()
        p = rec.get("plan") or {}
        print(label, tile_y, stages, "->", p.get("tile_y"), p.get("stages"), p.get("threads"), p.get("ctas"),
              round(rec.get("gcell_s", 0), 1), rec.get("same_bits_as_default"), rec.get("error", ""), flush=True)


if __name__ == "__main__":
  main()
