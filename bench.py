#!/usr/bin/env python
"""Headline benchmark: FDTD Gcell-updates/s on BASELINE.json's configuration.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

A "step" is ONE engine call: the full time-stepping of the workload (cfg2 at N=1: 90-degree
bend, 256x256x128 total grid, 20 000 FDTD steps, fp32) through the C ABI.
  value  : whole-job Gcell-updates/s, inputs resident in HBM, CUDA events, max over ranks.
  e2e    : same metric through `pjz_b200.fdtdz_jax.fdtdz` with HOST (pinned) buffers --
           host->device copies of every input and device->host copy of the snapshots inside
           the timed region.
  roofline: 60 B (fp32) per cell-update against the measured HBM copy peak.
For N > 1 (torchrun) every rank runs an independent engine call of the same workload with its
own source port -- the port/frequency batch axis of SURVEY.md 8(e); no data-path collective
("scaling": "weak").
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdtd_cell_updates_per_s"
UNIT = "Gcell-updates/s"
BYTES_PER_CELL = {False: 60.0, True: 30.0}


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="bend", choices=["bend", "waveguide", "demux", "coupler"])
  ap.add_argument("--tt", type=int, default=0, help="override the number of FDTD steps")
  ap.add_argument("--kernel", default="auto", choices=["auto", "twopass", "systolic", "systolic_async",
                                                       "systolic_lean"])
  ap.add_argument("--tile-y", type=int, default=0)
  ap.add_argument("--stages", type=int, default=0)
  ap.add_argument("--threads", type=int, default=0)
  ap.add_argument("--prefetch", type=int, default=0)
  ap.add_argument("--cols", type=int, default=0)
  ap.add_argument("--reduced", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-cpu", action="store_true")
  ap.add_argument("--cpu-seconds", type=float, default=12.0)
  ap.add_argument("--no-ab", action="store_true", help="skip the same-box A/B of the L2 prefetch setting")
  ap.add_argument("--no-reduced", action="store_true",
                  help="skip the reduced-precision sub-record")
  ap.add_argument("--no-decomp", action="store_true",
                  help="skip the domain-decomposed (cfg5 metalens, y-slabs) record")
  ap.add_argument("--decomp-tt", type=int, default=2000)
  ap.add_argument("--decomp-cols", type=int, default=512, help="owned y-columns per GPU")
  ap.add_argument("--decomp-x", type=int, default=4096)
  ap.add_argument("--decomp-transport", default="auto", choices=["auto", "p2p", "nccl"])
  return ap.parse_args()


def make_workload(args, rank):
  """Engine kwargs (NumPy, host) for the chosen BASELINE configuration."""
  from pjz_b200 import _field as glue
  from pjz_b200 import workloads as W
  if args.workload == "bend":
    eps, ports, params, omega = W.bend(reduced=args.reduced)
    name = "cfg2 90-degree bend 256x256x128 total grid, 20000 steps"
  elif args.workload == "waveguide":
    eps, ports, params, omega = W.straight_waveguide(reduced=args.reduced)
    name = "cfg1 straight waveguide 96x96x80 total grid, 4000 steps"
  elif args.workload == "demux":
    eps, ports, params, omega = W.demux(reduced=args.reduced)
    name = "cfg3 demux 512x512x128 total grid, 20000 steps, 4 frequencies"
  else:
    eps, ports, params, omega, _ = W.coupler(reduced=args.reduced)
    name = "cfg4 coupler 384x256x128 total grid, 20000 steps, one port per GPU"
  if args.tt:
    params = params._replace(tt=args.tt)
    name += f" (tt overridden to {args.tt})"
  lp = {"kernel": args.kernel}
  for k, v in (("tile_y", args.tile_y), ("stages", args.stages), ("threads", args.threads),
               ("prefetch", args.prefetch), ("cols", args.cols)):
    if v:
      lp[k] = v
  params = params._replace(launch_params=lp)
  axis, pos, _ = ports[rank % len(ports)]
  src = W.gaussian_port_source(eps, axis, pos)
  kw, _, _ = glue.engine_inputs(eps, src, omega, pos, params)
  host = {}
  for k, v in kw.items():
    host[k] = v.numpy() if hasattr(v, "numpy") else v
  X, Y = host["absorption_mask"].shape[1:]
  Z = host["pml_kappa"].shape[0]
  return host, (X, Y, Z), params.tt, name


ARRAYS = ("epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa",
          "pml_sigma", "pml_alpha")


class ClockSampler:
  """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    try:
      self.p = subprocess.Popen(
          ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
           "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
    except OSError:
      self.p = None

  def stop(self):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if self.p is None:
      return out
    self.p.terminate()
    try:
      self.p.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.p.kill()
    self.f.flush()
    self.f.seek(0)
    sm, mx, power, reasons = [], [], [], set()
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    for line in self.f.read().splitlines():
      c = [v.strip() for v in line.split(",")]
      if len(c) < 9:
        continue
      try:
        sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
      except ValueError:
        continue
      for n, v in zip(names, c[5:9]):
        if v.lower().startswith("active"):
          reasons.add(n)
    os.unlink(self.f.name)
    if sm:
      out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm),
                 power_w_max=max(power), reasons=sorted(reasons))
    return out


def measured_peak():
  try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  except Exception:
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
  """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
  try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
      return json.load(f)
  except Exception:
    return None


def line_config(name, dims, tt, world, working_set, need_flush):
  """`config` of the JSON line -- the SAME dict from both arms (the driver compares them): what is
  run, how it is spread over the devices, and how the L2 is kept cold between timed steps.  What
  only one arm has (the GPU plan) is a top-level key of that arm's line."""
  return {"workload": name, "grid": list(dims), "fdtd_steps": tt,
          "parallelism": "one engine run on one device" if world == 1 else
          f"port batch: {world} independent engine runs, one per device, no collective",
          "l2": (f"working set {working_set / 2**20:.0f} MiB: " +
                 ("smaller than 3x the L2, 512 MiB flush write between timed steps" if need_flush else
                  "larger than 3x the L2, no flush needed"))}


def working_set_bytes(dims, reduced):
  el = 2 if reduced else 4
  return dims[0] * dims[1] * dims[2] * (15 * el + 6 * el)


B200_L2_BYTES = 126 << 20              # the reference arm has no device to ask


def host_threads():
  """All host threads the process may use (torchrun pins OMP_NUM_THREADS=1; ignore that)."""
  try:
    return max(1, len(os.sched_getaffinity(0)))
  except AttributeError:
    return os.cpu_count() or 1


def cpu_sample(host, dims, seconds, reduced):
  """Oracle port (C, OpenMP, all host threads) on a bounded sample of the SAME workload:
  the same domain and inputs, the first n time steps; set-up time removed by differencing."""
  from oracle import fdtd_c
  cells = dims[0] * dims[1] * dims[2]
  kw = dict(host)
  kw["launch_params"] = None
  nthr = host_threads()

  def timed(n):
    t0 = time.perf_counter()
    fdtd_c.fdtdz(**kw, steps_override=n, want_output=False, nthreads=nthr)
    return time.perf_counter() - t0

  timed(0)
  t_setup = timed(0)
  n1 = 4
  t1 = timed(n1)
  rate = cells * n1 / max(t1 - t_setup, 1e-6)
  n2 = int(max(8, min(host["source_waveform"].shape[0], seconds * rate / cells)))
  t2 = timed(n2)
  value = cells * n2 / max(t2 - t_setup, 1e-6) / 1e9
  return {"value": value, "unit": UNIT, "cores": nthr, "kind": "port",
          "sample": f"same workload, first {n2} of {host['source_waveform'].shape[0]} FDTD steps "
                    f"({t2 - t_setup:.1f} s of CPU work, set-up excluded), fp"
                    f"{'16-storage' if reduced else '32'} C oracle with OpenMP"}


def run_reference(args, rank, world):
  """CPU arm: the oracle port (the engine's source is not in /root/reference -- SURVEY.md 8c),
  all host threads, same config/metric; each step is a bounded sample of the workload."""
  if rank != 0:
    return
  from oracle import fdtd_c
  host, dims, tt, name = make_workload(args, 0)
  cells = dims[0] * dims[1] * dims[2]
  kw = dict(host)
  kw["launch_params"] = None
  nthr = host_threads()

  def timed(n):
    t0 = time.perf_counter()
    fdtd_c.fdtdz(**kw, steps_override=n, want_output=False, nthreads=nthr)
    return time.perf_counter() - t0

  timed(0)
  t_setup = min(timed(0), timed(0))
  n = 100
  for _ in range(args.warmup):
    timed(4)
  dt = sum(timed(n) for _ in range(args.steps)) - args.steps * t_setup
  value = cells * n * args.steps / max(dt, 1e-9) / 1e9
  sample = (f"each step = first {n} of {tt} FDTD steps of the same workload, all {nthr} host "
            f"threads; set-up ({t_setup:.2f} s per call) excluded")
  print(json.dumps({
      "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f16-storage/f32-math" if args.reduced else "f32", "data": "synthetic",
      "config": line_config(name, dims, tt, world, working_set_bytes(dims, args.reduced),
                            working_set_bytes(dims, args.reduced) < 3 * B200_L2_BYTES),
      "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port",
                       "sample": sample},
      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def reduced_records(args, fdtdz_jax, torch):
  """Reduced precision (pjz's default mode): cfg2 with fp16 storage, and pjz's own default height
  (128 - sum(pml_widths) = 96 z-cells, /root/reference/src/pjz/_field.py:52-58)."""
  reduced = []
  for label, zz in (("cfg2 grid with fp16 storage (256x256x128)", 128),
                    ("pjz default height, fp16 storage (256x256x96)", 96)):
    from pjz_b200 import _field as glue
    from pjz_b200 import workloads as W
    eps, ports, params, omega = W.bend(total=(256, 256, zz), reduced=True)
    if args.tt:
      params = params._replace(tt=args.tt)
    axis, pos, _ = ports[0]
    kw, _, _ = glue.engine_inputs(eps, W.gaussian_port_source(eps, axis, pos), omega, pos, params)
    kw = {k: (v.cuda() if hasattr(v, "cuda") else v) for k, v in kw.items()}
    rinfo = fdtdz_jax.plan_info(**kw)
    fdtdz_jax.fdtdz(**kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
      r = fdtdz_jax.fdtdz(**kw)
    b.record()
    torch.cuda.synchronize()
    rcells = 256 * 256 * zz
    rv = rcells * params.tt * 2 / (a.elapsed_time(b) / 1e3) / 1e9
    assert bool(torch.isfinite(r).all())
    reduced.append({"value": rv, "unit": UNIT, "frac": rv * BYTES_PER_CELL[True] / measured_peak()[0],
                    "bytes_per_cell_update": BYTES_PER_CELL[True],
                    "config": {"workload": label, "fdtd_steps": params.tt, "plan": rinfo}})
  return reduced


# ---- the sharded-domain configuration (BASELINE.json config 5) ------------------------------------

def _metalens_slab(total, rank, world, ghost, tt, device, seed_wave=None):
  """Engine kwargs of ONE rank's y-slab of the cfg5 metalens (`ghost` ghost columns per side),
  built on the device: (local kwargs, owned columns, crop) as pjz_b200._decomp.local_problem_y
  would return them from the global arrays, which are never materialised."""
  import torch
  from pjz_b200 import _field as glue
  from pjz_b200 import workloads as W
  from pjz_b200._decomp import slab_bounds
  X, Y, Z = total
  pad, pml = 32, (16, 16)
  yy, zz = Y - 2 * pad, Z - sum(pml)
  y0, y1 = slab_bounds(Y, world, rank)
  cols = np.arange(y0 - ghost, y1 + ghost) % Y
  eps = W.metalens_columns(total, np.clip(cols - pad, 0, yy - 1), device, pad=pad, pml=pml)
  mask = glue._absorption_mask(X, Y, pad, 1e-4)
  t = np.arange(tt)
  rs = glue._ramped_sin(np.array([W.OMEGA0]), 4.0, 4.0, 0.5, tt)
  wf = np.stack([rs.imag[:, 0] if rs.ndim == 2 else rs.imag, rs.real[:, 0] if rs.ndim == 2 else rs.real], -1)
  # z-plane plane-wave source in the substrate (quadrature pair on Ex), as field() builds it
  src = torch.zeros((2, 2, X, len(cols), 1), dtype=torch.float32, device=device)
  src[1, 0] = 0.01
  loc = dict(
      epsilon=eps, dt=0.5, source_field=src, source_waveform=wf.astype(np.float32),
      source_position=pml[0] + zz // 6, absorption_mask=np.ascontiguousarray(mask[:, :, cols]),
      pml_kappa=np.ones((Z, 2), np.float32), pml_sigma=glue._pml_sigma(pml, Z, 0.5, 1.3),
      pml_alpha=np.full((Z, 2), 0.05, np.float32), pml_widths=pml,
      output_steps=(tt - 1, tt, 1), use_reduced_precision=False, launch_params=None,
      offset=(pad, 0, pml[0]))
  g0, g1 = max(pad, y0), min(pad + yy, y1)
  crop = (g0 - y0 + ghost, g1 - y0 + ghost, g0 - pad, g1 - pad) if g1 > g0 else None
  return loc, y1 - y0, crop


def _decomp_check(rank, world, transport):
  """Small domain, every rank: the N-rank decomposed result against the one-call engine."""
  import torch
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import fdtdz_decomposed_p2p, fdtdz_decomposed_y
  from tests.problems import random_problem
  kw = random_problem(domain=(48, 32 * world, 128), axis=2, pml=(16, 16), tt=60, seed=321,
                      output_steps=(20, 60, 13), absorb_pad=6, absorb_coeff=1e-3)
  if transport == "p2p":
    got = fdtdz_decomposed_p2p(**kw)
  else:
    got = fdtdz_decomposed_y(**kw, ghost=None if world > 1 else 8)
  dev = dict(kw)
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  want = fdtdz_jax.fdtdz(**dev)
  same = bool(torch.equal(got, want)) and bool(torch.isfinite(want).all()) and float(want.abs().max()) > 0
  return "bit-exact" if same else "MISMATCH"


def run_decomp(args, rank, world, local):
  """BASELINE.json config 5: the metalens domain cut into y-slabs, one per GPU, weak scaling
  (`--decomp-cols` owned columns x `--decomp-x` planes x 128 per GPU: 4096x4096x128 at 8 GPUs),
  `--decomp-tt` steps.  Transport "p2p": halo exchange from inside the persistent kernel through
  peer-mapped memory (one launch per GPU); "nccl": ghost zones + one packed send/recv per G steps.
  Also times the SAME slab wrapped onto itself on one GPU (no neighbour) for the efficiency."""
  import torch
  import torch.distributed as dist
  from pjz_b200._decomp import P2PSlabRun, YSlabRun, choose_ghost
  dev = torch.device("cuda", local)
  X, cols, Z, tt = args.decomp_x, args.decomp_cols, 128, args.decomp_tt
  total = (X, cols * world, Z)
  cells = X * cols * Z

  def agree(ok):
    f = torch.tensor([int(ok)], device=dev)
    if world > 1:
      dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(f.item())

  def build(transport, r, w, steps):
    if transport == "p2p":
      loc, nloc, crop = _metalens_slab((X, cols * w, Z), r, w, 1, steps, dev)
      return P2PSlabRun(None, local=(loc, nloc, crop), solo=(w != world))
    shapes = dict(absorption_mask=np.broadcast_to(np.float32(0), (3, X, cols * w)),
                  epsilon=np.broadcast_to(np.float32(0), (3, X - 64, cols * w - 64, Z - 32)),
                  source_field=np.broadcast_to(np.float32(0), (2, 2, X, cols * w, 1)),
                  pml_kappa=np.ones((Z, 2), np.float32), source_position=32, offset=(32, 32, 16),
                  dt=0.5, source_waveform=np.zeros((steps, 2), np.float32),
                  pml_sigma=np.zeros((Z, 2), np.float32), pml_alpha=np.zeros((Z, 2), np.float32),
                  pml_widths=(16, 16), output_steps=(steps - 1, steps, 1),
                  use_reduced_precision=False, launch_params=None)
    G = choose_ghost(shapes, w)
    loc, nloc, crop = _metalens_slab((X, cols * w, Z), r, w, G, steps, dev)
    kw = dict(shapes, epsilon=np.zeros((3, X - 64, cols * w - 64, 0), np.float32))
    run = YSlabRun(kw, ghost=G, local=(loc, nloc, crop), solo=(w != world))
    run.tt = steps
    return run

  note = None
  want = args.decomp_transport
  transports = []
  if want in ("auto", "p2p"):
    try:
      probe = build("p2p", rank, world, 8)
      probe.run(); torch.cuda.synchronize(); probe.close()
      ok = True
    except Exception as e:                                   # noqa: BLE001
      ok, note = False, f"p2p unavailable: {str(e)[:160]}"
    if agree(ok):
      transports.append("p2p")
    elif want == "p2p":
      raise RuntimeError(note or "p2p transport failed on another rank")
  if want in ("auto", "nccl"):
    transports.append("nccl")

  def timed(transport, w_, r_):
    warm = build(transport, r_, w_, 64)
    warm.run(); torch.cuda.synchronize(); warm.close(); del warm
    run = build(transport, r_, w_, tt)
    run.time_exchange = True
    torch.cuda.synchronize()
    if w_ > 1:
      dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run.run(); b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    lo, hi, snaps = run.local_snapshots()
    chk = float(snaps.double().abs().sum().item())
    info = (run.G, run.slab.kernel, run.slab.stages, getattr(run, "exchange_ms", None))
    run.close()
    return ms, chk, info

  records = {}
  for transport in transports:
    check = _decomp_check(rank, world, transport)
    ms_n, chk, (ghost, kernel, stages, xms) = timed(transport, world, rank)
    if world > 1:
      dist.all_reduce(ms_n, op=dist.ReduceOp.MAX)
    value_n = world * cells * tt / (float(ms_n.item()) / 1e3) / 1e9
    # the same slab on ONE GPU, wrapped onto itself (every rank times its own; rank 0's is quoted)
    if world > 1:
      ms_1, _, _ = timed(transport, 1, 0)
      value_1 = cells * tt / (float(ms_1.item()) / 1e3) / 1e9
    else:
      value_1 = value_n
    assert np.isfinite(chk) and chk > 0, "decomposed run produced an empty/non-finite field"
    records[transport] = {
        "value": value_n, "unit": UNIT, "ms_per_run": float(ms_n.item()),
        "value_1gpu_same_slab": value_1, "efficiency_vs_1gpu": value_n / (world * value_1),
        "transport": ("halo exchange inside the persistent launch: a courier CTA copies the edge "
                      "columns into the neighbours' ghost columns through peer-mapped memory (CUDA "
                      "IPC over NVLink) and forwards the progress counters with st.release.sys; one "
                      "launch per GPU, nothing exchanged by the host"
                      if transport == "p2p" else
                      f"ghost zones: NCCL send/recv of {ghost} ghost columns per side every {ghost} "
                      f"steps, one persistent launch per {ghost} steps ({(cols + 2 * ghost) / cols:.3f}x "
                      f"redundant columns)"),
        "ghost": ghost, "kernel": f"{kernel}, {stages} stages",
        "exchange_ms": xms if transport == "nccl" else None,
        "decomp_check": check}
  best = max(records, key=lambda k: records[k]["value"])
  out = dict(records[best])
  out.update({
      "n_gpus": world, "scaling": "weak", "selected": best,
      "config": {"workload": f"cfg5 metalens {X}x{cols * world}x{Z} total grid, y-slabs of {cols} "
                             f"columns per GPU, {tt} steps, fp32", "grid": [X, cols * world, Z],
                 "fdtd_steps": tt, "kernel": out["kernel"]},
      "all_transports": {"in_kernel_p2p": records.get("p2p"), "nccl_ghost_zones": records.get("nccl")},
      "note": note})
  return out


def main():
  args = parse()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if args.impl == "reference":
    run_reference(args, rank, world)
    return

  import torch
  import torch.distributed as dist
  from pjz_b200 import fdtdz_jax
  fdtdz_jax.lib()                      # fail loudly if the CUDA library is missing
  assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
  torch.cuda.set_device(local)
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

  host, dims, tt, name = make_workload(args, rank)
  cells = dims[0] * dims[1] * dims[2]
  dev = dict(host)
  for k in ARRAYS:
    dev[k] = torch.from_numpy(np.ascontiguousarray(host[k], np.float32)).cuda()
  info = fdtdz_jax.plan_info(**dev)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- value: inputs resident in HBM ----------------------------------------------------------
  # Cold L2 between timed steps: the default workload's working set (fields x2 + coefficients)
  # is several times the 126 MB L2; smaller workloads get an explicit flush (a 512 MB write)
  # between steps, outside the per-step CUDA events.
  working_set = working_set_bytes(dims, args.reduced)
  l2_bytes = torch.cuda.get_device_properties(local).L2_cache_size
  need_flush = working_set < 3 * l2_bytes
  flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if need_flush else None
  for _ in range(args.warmup):
    out = fdtdz_jax.fdtdz(**dev)
  barrier()
  sampler = ClockSampler(local) if rank == 0 else None
  events = []
  for _ in range(args.steps):
    if need_flush:
      flush_buf.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fdtdz_jax.fdtdz(**dev)
    b.record()
    events.append((a, b))
  barrier()
  ev_ms = sum(a.elapsed_time(b) for a, b in events)
  ms = torch.tensor([ev_ms], device="cuda")
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms.item())
  clocks = sampler.stop() if sampler else None
  checksum = float(out.double().abs().sum().item())
  assert np.isfinite(checksum) and checksum > 0, "engine produced an empty/non-finite field"
  value = world * cells * tt * args.steps / (ms / 1e3) / 1e9

  # ---- this box: device copy bandwidth, the way MEASURED_PEAKS.json's hbm_gbs was taken (the pool is
  # bimodal for the fp32 flagship; this says what kind of box the line comes from)
  box = None
  if rank == 0:
    src_t = torch.empty(1 << 29, dtype=torch.bfloat16, device="cuda")
    dst_t = torch.empty_like(src_t)
    best = 0.0
    for _ in range(6):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(); dst_t.copy_(src_t); b.record()
      torch.cuda.synchronize()
      best = max(best, 2 * src_t.numel() * 2 / (a.elapsed_time(b) / 1e3) / 1e9)
    box = {"copy_gbs_now": best, "how": "torch b.copy_(a) over 512 Mi bf16 elements, best of 6"}
    del src_t, dst_t

  # ---- same-box reference: the one setting round 2 changed in the flagship kernel's plan (the L2
  # prefetch of the service warp, 6 planes ahead in round 1, off now), both ways on a 4 000-step cut
  # of the same workload -- the pool's boxes differ by 25 % on this kernel, so a
  # number from another box says little
  same_box = None
  if rank == 0 and not args.no_ab and args.workload == "bend" and not args.reduced:
    try:
      short = dict(dev)
      short["source_waveform"] = dev["source_waveform"][:4000].contiguous()
      short["output_steps"] = (3997, 4000, 1)
      res_ab = {}
      for label, pf in (("l2_prefetch_off_this_build", None), ("l2_prefetch_6_planes_round1_setting", "6")):
        old_pf = os.environ.pop("B200FDTD_PF_AHEAD", None)
        if pf is not None:
          os.environ["B200FDTD_PF_AHEAD"] = pf
        try:
          fdtdz_jax.fdtdz(**short)
          torch.cuda.synchronize()
          a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record()
          for _ in range(3):
            fdtdz_jax.fdtdz(**short)
          b.record()
          torch.cuda.synchronize()
          res_ab[label] = cells * 4000 * 3 / (a.elapsed_time(b) / 1e3) / 1e9
        finally:
          os.environ.pop("B200FDTD_PF_AHEAD", None)
          if old_pf is not None:
            os.environ["B200FDTD_PF_AHEAD"] = old_pf
      same_box = dict(res_ab, unit=UNIT, workload="first 4000 steps of the same workload, 3 calls each")
    except Exception as e:                                   # noqa: BLE001
      same_box = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

  # ---- e2e: HOST pinned buffers through the public call (copies inside the timed region) ----
  e2e = None
  if not args.no_e2e:
    pinned = dict(host)
    for k in ARRAYS:
      t = torch.from_numpy(np.ascontiguousarray(host[k], np.float32)).pin_memory()
      pinned[k] = t.numpy()
    h2d = sum(pinned[k].nbytes for k in ARRAYS)
    res = fdtdz_jax.fdtdz(**pinned)    # warm-up (also first-touch of the host path)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
      res = fdtdz_jax.fdtdz(**pinned)  # NumPy in -> b200fdtd_run_host -> NumPy out (synchronous)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
      dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": world * cells * tt * args.steps / float(t_e2e.item()) / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(res.nbytes)}
    assert np.isfinite(res).all()

  # ---- reduced precision (pjz's default mode): same workload with fp16 storage, and pjz's own
  # default height (128 - sum(pml_widths) = 96 z-cells, /root/reference/src/pjz/_field.py:52-58)
  reduced = None
  if not args.reduced and not args.no_reduced and args.workload == "bend" and world == 1:
    try:                                           # the headline must not be lost to this record
      reduced = reduced_records(args, fdtdz_jax, torch)
    except Exception as e:                         # noqa: BLE001
      reduced = [{"error": f"{type(e).__name__}: {str(e)[:300]}"}]

  # ---- the sharded-domain record (all ranks take part) ----------------------------------------------
  decomp = None
  if not args.no_decomp:
    del out, dev
    torch.cuda.empty_cache()
    if world == 1:
      try:                                         # the headline must not be lost to this record
        decomp = run_decomp(args, rank, world, local)
      except Exception as e:                       # noqa: BLE001
        decomp = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    else:                                          # (collective: a rank cannot fail alone)
      decomp = run_decomp(args, rank, world, local)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel -----------------------------------------------------------
  peak, peak_src = measured_peak()
  bpc = BYTES_PER_CELL[bool(args.reduced)]
  per_gpu = value / world
  achieved = per_gpu * bpc                       # GB/s of ALGORITHMIC traffic
  traffic = ncu_traffic()
  if info["kernel"].startswith("systolic"):
    kname = {"systolic_lean": "lean_kernel",
             "systolic_async": "systolic2_kernel"}.get(info["kernel"], "systolic_kernel")
    if info["kernel"] == "systolic_lean" and (args.reduced or dims[2] <= 64):
      kname = "lean16_kernel"                    # sub-warp variant: columns of <= 16 vectors
    dominant = kname + " (1 launch per engine call; duration = call time incl. 3 prep kernels)"
    launches = (3 + 2) * args.steps
  else:
    dominant = "twopass_h_kernel + twopass_e_kernel (2 launches per FDTD step)"
    launches = (3 + 2 * tt) * args.steps
  per_launch_updates = cells * (tt if info["kernel"].startswith("systolic") else 0.5)
  # ncu DRAM bytes are only quoted for the kernel and configuration they were captured on
  traffic = (traffic or {}).get("fp16" if args.reduced else "fp32")
  have_traffic = (traffic is not None and traffic.get("plan_kernel") == info["kernel"] and
                  args.workload == "bend")
  roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
              "frac": achieved / peak,
              "traffic": traffic["dram_bytes_per_cell_update"] * per_launch_updates
              if have_traffic else None,
              "traffic_source": traffic["from"] if have_traffic else None,
              "kernel": dominant, "peak_source": peak_src,
              "algorithmic_bytes_per_launch": bpc * per_launch_updates,
              "bytes_per_cell_update": bpc,
              # `frac` is a fraction of the SINGLE-PASS bound (60 / 30 algorithmic bytes per
              # cell-update); the kernel is temporally blocked through L2, so its real HBM load is
              # dram_frac = ncu DRAM bytes per cell-update x measured rate / peak -- HBM is not what
              # limits it (issue latency is: issue_active_pct)
              "dram_frac": traffic["dram_bytes_per_cell_update"] * per_gpu / peak
              if have_traffic else None,
              "l2_throughput_pct_ncu": traffic["l2_throughput_pct"] if have_traffic else None,
              "issue_active_pct_ncu": traffic["issue_active_pct"] if have_traffic else None}

  cpu = None
  if world == 1 and not args.no_cpu:
    cpu = cpu_sample(host, dims, args.cpu_seconds, args.reduced)

  line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None,
      "dtype": "f16-storage/f32-math" if args.reduced else "f32", "data": "synthetic",
      "config": line_config(name, dims, tt, world, working_set, need_flush),
      "plan": info, "l2_bytes": int(l2_bytes),
      "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
      "clocks": clocks, "box": box, "same_box_ab": same_box, "reduced_precision": reduced, "decomp": decomp,
  }
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
