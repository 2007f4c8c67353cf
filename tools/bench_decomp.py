"""Strong-scaling measurement of the x-slab domain decomposition (pjz_b200/_decomp.py).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        tools/bench_decomp.py --grid 2048 1024 128 --tt 200

One domain, N ranks; each rank owns X/N planes and exchanges one H and one E face per step.
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
  # metalens-like stack built cheaply on the host: substrate / pillar layer / air
  eps = np.full((3, xx, yy, zz), 1.0, np.float32)
  eps[..., : zz // 3] = 2.25
  rng = np.random.default_rng(1)
  pillars = (rng.random((xx // 16 + 1, yy // 16 + 1)) > 0.5).astype(np.float32)
  layer = np.kron(pillars, np.ones((16, 16), np.float32))[:xx, :yy]
  eps[..., zz // 3: zz // 3 + 24] = (1.0 + 11.25 * layer)[None, :, :, None]
  t = np.arange(args.tt)
  wf = np.stack([np.sin(2 * np.pi / 37 * 0.5 * t), np.zeros_like(t, dtype=np.float64)], -1)
  kw = dict(
      epsilon=eps, dt=0.5,
      source_field=np.ones((2, 2, X, Y, 1), np.float32) * 0.01,
      source_waveform=wf.astype(np.float32), source_position=16 + 8,
      absorption_mask=glue._absorption_mask(X, Y, pad, 1e-4),
      pml_kappa=np.ones((Z, 2), np.float32), pml_sigma=glue._pml_sigma(pml, Z, 0.5, 1.3),
      pml_alpha=np.zeros((Z, 2), np.float32), pml_widths=pml,
      output_steps=(args.tt - 1, args.tt, 1), use_reduced_precision=False, launch_params=None,
      offset=(pad, pad, pml[0]))
  from pjz_b200._decomp import DecomposedRun
  for _ in range(args.warmup):
    small = dict(kw)
    small["source_waveform"] = kw["source_waveform"][:4]
    small["output_steps"] = (3, 4, 1)
    fdtdz_decomposed(**small, gather=False)
  run = DecomposedRun(kw)                      # set-up (slab inputs, H2D, coefficients): untimed
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
        "config": {"workload": "metalens-like stack, x-slab decomposition, halo exchange per half-step",
                   "grid": [X, Y, Z], "fdtd_steps": args.tt, "kernel": "twopass (session API)"},
        "ms_per_fdtd_step": float(ms) / args.tt, "checksum": float(chk)}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
