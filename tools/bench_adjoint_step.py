"""BASELINE cfg3: one adjoint optimisation step of the wavelength demux on ONE GPU, end to end.

    python tools/bench_adjoint_step.py [--tt 20000]

mode_gpu (2 ports x 4 frequencies) -> scatter() = 2 engine runs (input port = "forward",
output port = "adjoint" by reciprocity, /root/reference/src/pjz/_field.py:346-398), 9 snapshots
each -> phasor projection -> overlaps -> loss.backward() through the fused product-reduce
kernel.  Prints one JSON line: wall seconds per stage, Gcell-updates/s of the whole step.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--tt", type=int, default=20000)
  ap.add_argument("--fuse-projection", action="store_true")
  args = ap.parse_args()
  from pjz_b200 import scatter
  from pjz_b200 import workloads as W
  from pjz_b200._mode_gpu import mode_gpu
  eps, ports, params, omega = W.demux()
  params = params._replace(tt=args.tt)
  X, Y, Z = (eps.shape[1] + 2 * params.absorption_padding, eps.shape[2] + 2 * params.absorption_padding,
             params.domain_zz)
  e = torch.from_numpy(np.ascontiguousarray(eps)).cuda()

  def step():
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    modes, betas = [], []
    for axis, pos, _ in ports:
      sl = [slice(None)] * 4
      sl["xyz".index(axis) + 1] = slice(pos, pos + 1)
      b, m, _, _ = mode_gpu(e[tuple(sl)], omega, 1)
      modes.append(m[..., 0]); betas.append(b[:, 0].cpu().numpy())
    torch.cuda.synchronize(); t["modes_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    x = e.clone().requires_grad_(True)
    sv = scatter(x, omega, modes, betas, [p[1] for p in ports], [True, False], params,
                 fuse_projection=args.fuse_projection)
    loss = -(sv[0][1].abs() ** 2).sum()                     # maximise transmission in -> out
    torch.cuda.synchronize(); t["forward_s"] = time.perf_counter() - t0; t0 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize(); t["backward_s"] = time.perf_counter() - t0
    t["loss"] = float(loss)
    t["grad_norm"] = float(x.grad.norm())
    return t

  small = params
  params = params._replace(tt=2500)
  step()                                                    # warm-up (library load, cuSOLVER, ...)
  params = small
  t = step()
  total = t["modes_s"] + t["forward_s"] + t["backward_s"]
  cells = X * Y * Z
  print(json.dumps({"workload": f"cfg3 demux adjoint step {X}x{Y}x{Z}, {args.tt} steps, 2 engine runs, "
                    f"{omega.shape[0]} frequencies", **t, "total_s": total,
                    "gcell_updates_per_s_whole_step": 2 * cells * args.tt / total / 1e9,
                    "peak_mem_gib": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
  main()
