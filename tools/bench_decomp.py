"""Strong-scaling measurement of the domain decompositions (pjz_b200/_decomp.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        tools/bench_decomp.py --grid 2048 1024 128 --tt 200 [--mode y --ghost 14]

One domain, N ranks.  --mode x: each rank owns X/N planes and exchanges one H and one E face per
step (per-step kernels).  --mode y: each rank owns Y/N columns plus `ghost` ghost columns per
side and advances `ghost` steps per exchange with ONE persistent systolic launch.
Prints one JSON line (rank 0): whole-domain Gcell-updates/s, max over ranks, CUDA events.
"""

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--grid", type=int, nargs=3, default=[2048, 1024, 128])
  ap.add_argument("--tt", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=1)
  ap.add_argument("--mode", default="x", choices=["x", "y"])
  ap.add_argument("--ghost", type=int, default=0, help="0: a multiple of the pipeline depth >= 8")
  args = ap.parse_args()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  from pjz_b200 import _field as glue
  from pjz_b200._decomp import fdtdz_decomposed
  X, Y, Z = args.grid
  pad, pml = 32, (16, 16)
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  rng = np.random.default_rng(1)
  pillars = (rng.random((xx // 16 + 1, yy // 16 + 1)) > 0.5).astype(np.float32)

  def stack(ycols):
    """Metalens-like stack (substrate / pillar layer / air) for sub-volume columns `ycols`."""
    eps = np.full((3, xx, len(ycols), zz), 1.0, np.float32)
    eps[..., : zz // 3] = 2.25
    layer = pillars[np.arange(xx)[:, None] // 16, np.asarray(ycols)[None, :] // 16]
    eps[..., zz // 3: zz // 3 + 24] = (1.0 + 11.25 * layer)[None, :, :, None]
    return eps

  t = np.arange(args.tt)
  wf = np.stack([np.sin(2 * np.pi / 37 * 0.5 * t), np.zeros_like(t, dtype=np.float64)], -1)
  mask = glue._absorption_mask(X, Y, pad, 1e-4)
  common = dict(
      dt=0.5, source_waveform=wf.astype(np.float32), source_position=16 + 8,
      pml_kappa=np.ones((Z, 2), np.float32), pml_sigma=glue._pml_sigma(pml, Z, 0.5, 1.3),
      pml_alpha=np.zeros((Z, 2), np.float32), pml_widths=pml,
      output_steps=(args.tt - 1, args.tt, 1), use_reduced_precision=False, launch_params=None)
  from pjz_b200._decomp import DecomposedRun, YSlabRun, slab_bounds

  def small(d, nw):
    d = dict(d)
    d["source_waveform"] = d["source_waveform"][:nw]
    d["output_steps"] = (nw - 1, nw, 1)
    return d

  if args.mode == "x":
    kw = dict(common, epsilon=stack(np.arange(yy)), absorption_mask=mask,
              source_field=np.ones((2, 2, X, Y, 1), np.float32) * 0.01, offset=(pad, pad, pml[0]))
    for _ in range(args.warmup):
      fdtdz_decomposed(**small(kw, 4), gather=False)
    run = DecomposedRun(kw)                    # set-up (slab inputs, H2D, coefficients): untimed
  else:
    # every rank builds only ITS slab (ghost columns included): the global permittivity of the
    # BASELINE metalens (4096x4096x128) would not fit the host memory of 8 rank processes
    from pjz_b200._decomp import choose_ghost
    shapes = dict(common, epsilon=np.broadcast_to(np.float32(0), (3, xx, yy, zz)),
                  absorption_mask=np.broadcast_to(np.float32(0), (3, X, Y)),
                  source_field=np.broadcast_to(np.float32(0), (2, 2, X, Y, 1)),
                  offset=(pad, pad, pml[0]))
    G = args.ghost or choose_ghost(shapes, world)
    args.ghost = G
    y0, y1 = slab_bounds(Y, world, rank)
    cols = np.arange(y0 - G, y1 + G) % Y
    loc = dict(common, epsilon=stack(np.clip(cols - pad, 0, yy - 1)),
               absorption_mask=np.ascontiguousarray(mask[:, :, cols]),
               source_field=np.ones((2, 2, X, len(cols), 1), np.float32) * 0.01,
               offset=(pad, 0, pml[0]))
    g0, g1 = max(pad, y0), min(pad + yy, y1)
    crop = (g0 - y0 + G, g1 - y0 + G, g0 - pad, g1 - pad) if g1 > g0 else None
    kw = dict(common, epsilon=np.zeros((3, xx, yy, 0), np.float32))   # (shape only; never sliced)
    for _ in range(args.warmup):
      w = YSlabRun(kw, ghost=G, local=(small(loc, 2 * G), y1 - y0, crop))
      w.tt = 2 * G
      w.run()
      w.close()
    run = YSlabRun(kw, ghost=G, local=(loc, y1 - y0, crop))
    run.tt = args.tt
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  run.run()                                    # the time loop: 2 kernels + 2 face exchanges per step
  ev1.record()
  torch.cuda.synchronize()
  lo, hi, snaps = run.local_snapshots()
  ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
  chk = torch.tensor([float(snaps.double().abs().sum())], device="cuda")
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
  if rank == 0:
    cells = X * Y * Z
    print(json.dumps({
        "metric": "fdtd_cell_updates_per_s", "unit": "Gcell-updates/s",
        "value": cells * args.tt / (float(ms) / 1e3) / 1e9, "n_gpus": world, "scaling": "strong",
        "config": {"workload": "metalens-like stack, x-slab decomposition, halo exchange per half-step"
                   if args.mode == "x" else
                   f"metalens-like stack, y-slab decomposition, {args.ghost} ghost columns per side, "
                   f"one halo exchange per {args.ghost} steps",
                   "grid": [X, Y, Z], "fdtd_steps": args.tt,
                   "kernel": "twopass (session API)" if args.mode == "x" else
                   f"{run.slab.kernel} (session advance, {run.slab.stages} stages)"},
        "ms_per_fdtd_step": float(ms) / args.tt, "checksum": float(chk)}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
