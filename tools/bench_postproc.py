"""Measurement of the SURVEY.md 8(f1)/(f2) rows on one GPU (JSON lines on stdout).

f2  b200fdtd_adjoint_reduce vs the reference formula's N^2 volume temporaries (torch), GB/s of
    algorithmic traffic (N*ww*8 B in + 4 B out per voxel-component) against the HBM copy peak.
f1  field() on the cfg3 demux geometry with and without the fused projection (engine +
    projection time; the time stepping dominates, the fused path saves the snapshot round trip).
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ev_time(fn, reps=5):
  fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


def main():
  from pjz_b200 import _field as glue
  from pjz_b200 import workloads as W
  peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
  # ---- f2 ------------------------------------------------------------------------------------------
  for nports, ww, shape in [(4, 4, (448, 448, 96)), (8, 2, (448, 448, 96)), (2, 1, (192, 192, 96))]:
    g = torch.Generator(device="cuda").manual_seed(1)
    fields = [torch.view_as_complex(torch.randn((ww, 3) + shape + (2,), generator=g, device="cuda"))
              for _ in range(nports)]
    amps = [torch.view_as_complex(torch.randn((ww, 2), generator=g, device="cuda")) + 2 for _ in range(nports)]
    gm = [[torch.view_as_complex(torch.randn((ww, 2), generator=g, device="cuda")) for _ in range(nports)]
          for _ in range(nports)]
    nvox = 3 * shape[0] * shape[1] * shape[2]
    ms = ev_time(lambda: glue._scatter_bwd_fused(fields, amps, gm))

    def ref():
      grads = [[fi * fj / a[:, None, None, None, None] for fj in fields] for a, fi in zip(amps, fields)]
      return glue._scatter_bwd(grads, gm)
    ms_ref = ev_time(ref, reps=2)
    err = float((glue._scatter_bwd_fused(fields, amps, gm) - ref()).norm() / ref().norm())
    bytes_alg = nvox * (nports * ww * 8 + 4)
    print(json.dumps({"row": "f2 adjoint product-reduce", "nports": nports, "ww": ww, "volume": shape,
                      "ms": ms, "ms_torch_reference_formula": ms_ref, "speedup": ms_ref / ms,
                      "roofline": {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": peak,
                                   "unit": "GB/s", "frac": bytes_alg / ms / 1e6 / peak},
                      "rel_l2_vs_reference_formula": err}))
    del fields
    torch.cuda.empty_cache()
  # ---- a4 / f1: one-pass projection and fused overlaps -----------------------------------------------
  from pjz_b200 import fdtdz_jax
  for ww, shape in [(4, (448, 448, 96)), (1, (192, 192, 96)), (8, (192, 192, 96))]:
    g = torch.Generator(device="cuda").manual_seed(2)
    n_out = 2 * ww + 1
    snaps = torch.randn((n_out, 3) + shape, generator=g, device="cuda")
    Wm = torch.randn((2 * ww, n_out), generator=g, device="cuda")
    Wn = Wm.cpu().numpy()
    nvox = 3 * shape[0] * shape[1] * shape[2]
    ms = ev_time(lambda: fdtdz_jax.project(snaps, Wn))

    def ref():
      o = torch.einsum("ij,j...->i...", Wm, snaps)
      return torch.complex(o[:ww], o[ww:])
    ms_ref = ev_time(ref, reps=3)
    want = ref()
    err = float((fdtdz_jax.project(snaps, Wn) - want).abs().max() / want.abs().max())
    bytes_alg = nvox * (n_out * 4 + ww * 8)
    print(json.dumps({"row": "a4/f1 one-pass snapshot projection (b200fdtd_project)", "ww": ww,
                      "snapshots": n_out, "volume": shape, "ms": ms, "ms_torch_einsum_plus_complex": ms_ref,
                      "speedup": ms_ref / ms,
                      "roofline": {"bound": "hbm", "achieved": bytes_alg / ms / 1e6, "peak": peak,
                                   "unit": "GB/s", "frac": bytes_alg / ms / 1e6 / peak},
                      "max_abs_err_over_max": err}))
    del snaps, want
    torch.cuda.empty_cache()
  for nports in (2, 8):
    g = torch.Generator(device="cuda").manual_seed(3)
    ww, shape = 4, (448, 448, 96)
    if nports == 8:
      ww, shape = 1, (320, 192, 96)
    fields = [torch.view_as_complex(torch.randn((ww, 3) + shape + (2,), generator=g, device="cuda"))
              for _ in range(nports)]
    modes = [torch.randn((ww, 2, 1, shape[1], shape[2]), generator=g, device="cuda") for _ in range(nports)]
    betas = [np.full(ww, 0.33) for _ in range(nports)]
    pos = [6 + i for i in range(nports)]
    fwd = [i % 2 == 0 for i in range(nports)]
    ms = ev_time(lambda: glue._overlaps_fused(fields, modes, betas, pos, fwd))

    def ref_ov():
      amps = [glue._overlap(m, b, p, f, fl)[:, 0] for fl, m, b, p, f in zip(fields, modes, betas, pos, fwd)]
      return [[glue._overlap(m, b, p, f, fl)[:, 1] / a for m, b, p, f in zip(modes, betas, pos, fwd)]
              for a, fl in zip(amps, fields)]
    ms_ref = ev_time(ref_ov, reps=2)
    print(json.dumps({"row": "f1 port overlaps, all pairs in one launch (b200fdtd_overlaps)",
                      "nports": nports, "ww": ww, "volume": shape, "ms": ms,
                      "ms_eager_per_port_chains": ms_ref, "speedup": ms_ref / ms}))
    del fields
    torch.cuda.empty_cache()
  # ---- f3 ------------------------------------------------------------------------------------------
  from pjz_b200 import mode
  from pjz_b200._mode_gpu import mode_gpu
  for (uu, vv), ww, mm in [((30, 20), 1, 4), ((192, 96), 4, 2)]:
    eps1 = np.ones((3, 1, uu, vv), np.float32) * 2.25
    eps1[:, :, uu // 2 - 6:uu // 2 + 6, vv // 2 - 4:vv // 2 + 4] = 12.25
    om = np.linspace(2 * np.pi / 40, 2 * np.pi / 36, ww) if ww > 1 else np.array([2 * np.pi / 37])
    mode_gpu(eps1, om, mm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bg, _, eg, it = mode_gpu(eps1, om, mm)
    torch.cuda.synchronize()
    tg = time.perf_counter() - t0
    t0 = time.perf_counter()
    bh, _, _, _ = mode(eps1, om, mm)
    th = time.perf_counter() - t0
    print(json.dumps({"row": "f3 mode solver", "cross_section": [uu, vv], "ww": ww, "num_modes": mm,
                      "gpu_s": tg, "outer_iterations": it, "max_residual": float(eg.max()),
                      "host_arpack_s": th,
                      "max_rel_beta_diff_vs_host": float(np.max(np.abs(bg.cpu().numpy() - bh) / bh))}))
  # ---- f4 ------------------------------------------------------------------------------------------
  from oracle import render_numpy
  from pjz_b200._epsilon import render
  for ll, xx, yy, zz, m in [(3, 448, 448, 96, 2), (3, 64, 64, 32, 2)]:
    rng = np.random.default_rng(0)
    layers = torch.from_numpy(rng.uniform(1, 12.25, (ll, 2 * m * xx, 2 * m * yy)).astype(np.float32)).cuda()
    pos = np.linspace(zz / 3, 2 * zz / 3, ll - 1).astype(np.float32)
    gs = (np.arange(zz)[:, None] + np.array([[-0.5, 0]])).astype(np.float32)
    ge = (np.arange(zz)[:, None] + np.array([[0.5, 1]])).astype(np.float32)
    ms = ev_time(lambda: render(layers, pos, gs, ge, m))
    row = {"row": "f4 renderer", "layers": [ll, 2 * m * xx, 2 * m * yy], "out": [3, xx, yy, zz], "ms": ms,
           "bytes_algorithmic": 3 * layers.numel() * 4 + 3 * xx * yy * zz * 4}
    row["roofline"] = {"bound": "hbm", "achieved": row["bytes_algorithmic"] / ms / 1e6, "peak": peak,
                       "unit": "GB/s", "frac": row["bytes_algorithmic"] / ms / 1e6 / peak}
    if xx <= 64:
      t0 = time.perf_counter()
      want = render_numpy.render(layers.cpu().numpy(), pos, gs, ge, m)
      row["numpy_oracle_ms"] = (time.perf_counter() - t0) * 1e3
      got = render(layers, pos, gs, ge, m).cpu().numpy()
      row["max_rel_diff_vs_oracle"] = float(np.max(np.abs(got - want) / np.abs(want)))
    print(json.dumps(row))
  # ---- f1 ------------------------------------------------------------------------------------------
  eps, ports, params, omega = W.demux(reduced=False)
  params = params._replace(tt=int(os.environ.get("F1_TT", "3000")))
  axis, pos, _ = ports[0]
  src = W.gaussian_port_source(eps, axis, pos)
  e = torch.from_numpy(np.ascontiguousarray(eps)).cuda()
  res = {}
  for fuse in (False, True):
    glue.field(e, src, omega, pos, params, fuse_projection=fuse)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = glue.field(e, src, omega, pos, params, fuse_projection=fuse)
    torch.cuda.synchronize()
    res[fuse] = (time.perf_counter() - t0, out)
    torch.cuda.empty_cache()
  err = float((res[True][1] - res[False][1]).abs().max() / res[False][1].abs().max())
  print(json.dumps({"row": "f1 fused frequency projection", "workload": "cfg3 demux 512x512x128, "
                    f"{params.tt} steps, {omega.shape[0]} frequencies ({2 * omega.shape[0] + 1} snapshots)",
                    "field_s_snapshots_then_einsum": res[False][0], "field_s_fused": res[True][0],
                    "max_abs_diff_over_max": err}))


if __name__ == "__main__":
  main()
