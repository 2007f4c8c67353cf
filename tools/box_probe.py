"""What kind of box is this?  The pool is bimodal for the fp32 flagship kernel;
this prints what nvidia-smi says about the GPU next to a 4 000-step cfg2 cut and a copy bandwidth,
so that two boxes can be compared.  GPU only.

  python tools/box_probe.py > gpurun_out/box_probe.txt
"""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pjz_b200 import _field as glue                      # noqa: E402
from pjz_b200 import fdtdz_jax                           # noqa: E402
from pjz_b200 import workloads as W                      # noqa: E402

Q = ("name,uuid,vbios_version,driver_version,pci.bus_id,power.limit,power.max_limit,clocks.max.sm,"
     "clocks.max.mem,ecc.mode.current,temperature.gpu,temperature.memory")
L = "clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active"


def smi(query):
  return subprocess.run(["nvidia-smi", f"--query-gpu={query}", "--format=csv,noheader", "-i", "0"],
                        capture_output=True, text=True).stdout.strip()


def main():
  print("static:", Q)
  print(smi(Q))
  eps, ports, params, omega = W.bend()
  params = params._replace(tt=4000)
  axis, pos, _ = ports[0]
  kw, _, _ = glue.engine_inputs(eps, W.gaussian_port_source(eps, axis, pos), omega, pos, params)
  kw = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in kw.items()}
  fdtdz_jax.fdtdz(**kw)
  torch.cuda.synchronize()
  samples, stop = [], False

  def sample():
    while not stop:
      samples.append(smi(L))
      time.sleep(0.2)
  th = threading.Thread(target=sample)
  th.start()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(8):
    fdtdz_jax.fdtdz(**kw)
  b.record()
  torch.cuda.synchronize()
  stop = True
  th.join()
  print("cfg2 first 4000 steps x 8:", round(256 * 256 * 128 * 4000 * 8 / (a.elapsed_time(b) / 1e3) / 1e9, 1),
        "Gcell-updates/s")
  print("under load:", L)
  for s in samples[1:-1]:
    print(" ", s)
  src = torch.empty(1 << 29, dtype=torch.bfloat16, device="cuda")
  dst = torch.empty_like(src)
  best = 0.0
  for _ in range(6):
    a.record(); dst.copy_(src); b.record()
    torch.cuda.synchronize()
    best = max(best, 2 * src.numel() * 2 / (a.elapsed_time(b) / 1e3) / 1e9)
  print("copy GB/s:", round(best, 1))
  x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
  y = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
  (x @ y)
  torch.cuda.synchronize()
  a.record()
  for _ in range(20):
    x @ y
  b.record()
  torch.cuda.synchronize()
  print("bf16 GEMM TF/s:", round(20 * 2 * 8192 ** 3 / (a.elapsed_time(b) / 1e3) / 1e12, 1))


if __name__ == "__main__":
  main()
