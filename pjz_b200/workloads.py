"""Synthetic permittivity stacks for the BASELINE.json configurations.

Everything is deterministic.  Each builder returns ``(epsilon, ports, sim_params, omega)``
where ``epsilon`` is the user-domain ``(3, xx, yy, zz)`` float32 array that ``field()`` embeds
in the engine domain (absorber padding in x-y, z padding + PML), and ``ports`` is a list of
``(axis, position, is_fwd)``.  Sizes are quoted as TOTAL engine grid = user domain + padding.
"""

from __future__ import annotations

import numpy as np

from ._field import SimParams

EPS_SI, EPS_CLAD = 12.25, 2.25
OMEGA0 = 2 * np.pi / 37


def _yee(mask_fn, xx, yy, zz):
  """Sample a boolean occupancy function on the three staggered E sub-grids
  (/root/reference/src/pjz/_epsilon.py:13-24: Ex at (x+1/2,y,z), Ey at (x,y+1/2,z),
  Ez at (x,y,z+1/2))."""
  x, y, z = np.meshgrid(np.arange(xx, dtype=np.float32), np.arange(yy, dtype=np.float32),
                        np.arange(zz, dtype=np.float32), indexing="ij", sparse=True)
  shifts = ((0.5, 0, 0), (0, 0.5, 0), (0, 0, 0.5))
  eps = np.empty((3, xx, yy, zz), np.float32)
  for c, (sx, sy, sz) in enumerate(shifts):
    eps[c] = np.where(mask_fn(x + sx, y + sy, z + sz), EPS_SI, EPS_CLAD)
  return eps


def straight_waveguide(xx=64, yy=64, zz=64, width=12, thick=4, pad=16, pml=(8, 8), tt=4000,
                       dt=0.5, reduced=False):
  """cfg1: straight Si waveguide along x; engine grid (xx+2 pad, yy+2 pad, zz+sum(pml))."""
  yc, zc = yy / 2, zz / 2
  eps = _yee(lambda x, y, z: (np.abs(y - yc) < width / 2) & (np.abs(z - zc) < thick / 2) & (x > -1),
             xx, yy, zz)
  # absorber: quadratic profile peaking at 1e-3 * pad^2 = 0.26 per unit time (|S11| = 0.026 on the
  # C oracle; 4e-4 left 0.13 of reflection from the x ends); ports 35 cells apart
  params = SimParams(omega_range=(OMEGA0, OMEGA0), tt=tt, dt=dt, absorption_padding=pad,
                     absorption_coeff=1e-3, pml_widths=pml, use_reduced_precision=reduced,
                     domain_zz=zz + sum(pml))
  ports = [("x", 10, True), ("x", 45, False)]
  return eps, ports, params, np.array([OMEGA0])


def bend(total=(256, 256, 128), pad=32, pml=(16, 16), radius=64, width=12, thick=8, tt=20000,
         dt=0.5, reduced=False):
  """cfg2: 90-degree bend; TOTAL engine grid 256x256x128, fp32, 20k steps (BASELINE.json)."""
  X, Y, Z = total
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  zc = zz / 2
  y_in = yy / 2 - radius / 2          # input arm: along +x at y = y_in
  xb = xx / 2 - radius / 2            # bend starts at x = xb; centre of curvature (xb, y_in + R)
  cy = y_in + radius

  def core(x, y, z):
    in_z = np.abs(z - zc) < thick / 2
    arm_in = (x <= xb) & (np.abs(y - y_in) < width / 2)
    r = np.sqrt((x - xb) ** 2 + (y - cy) ** 2)
    arc = (x > xb) & (y < cy) & (np.abs(r - radius) < width / 2)
    arm_out = (y >= cy) & (np.abs(x - (xb + radius)) < width / 2)
    return in_z & (arm_in | arc | arm_out)

  eps = _yee(core, xx, yy, zz)
  params = SimParams(omega_range=(OMEGA0, OMEGA0), tt=tt, dt=dt, absorption_padding=pad,
                     absorption_coeff=1e-4, pml_widths=pml, use_reduced_precision=reduced,
                     domain_zz=Z)
  ports = [("x", 6, True), ("y", yy - 8, True)]
  return eps, ports, params, np.array([OMEGA0])


def demux(total=(512, 512, 128), pad=32, pml=(16, 16), design=256, width=12, thick=8, tt=20000,
          ww=4, dt=0.5, seed=0, reduced=False):
  """cfg3: wavelength demux: seeded binary-blurred design region, 1 in + 1 out port,
  ww frequencies over lambda in [36, 40] via output_steps."""
  X, Y, Z = total
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  rng = np.random.default_rng(seed)
  coarse = rng.random((design // 8 + 2, design // 8 + 2)) > 0.5
  blur = np.kron(coarse, np.ones((8, 8)))[:design, :design].astype(np.float32)
  k = np.ones(5, np.float32) / 5
  for ax in (0, 1):
    blur = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, blur)
  x0, y0 = (xx - design) // 2, (yy - design) // 2
  zc, yc = zz / 2, yy / 2
  eps = _yee(lambda x, y, z: (np.abs(z - zc) < thick / 2) & (np.abs(y - yc) < width / 2) &
             ((x < x0) | (x >= x0 + design)), xx, yy, zz)
  zsl = slice(int(zc - thick / 2), int(zc + thick / 2))
  eps[:, x0:x0 + design, y0:y0 + design, zsl] = (
      EPS_CLAD + (EPS_SI - EPS_CLAD) * blur)[None, :, :, None]
  wl = np.linspace(40, 36, ww)
  omega = 2 * np.pi / wl
  params = SimParams(omega_range=(float(omega.min()), float(omega.max())), tt=tt, dt=dt,
                     absorption_padding=pad, absorption_coeff=1e-4, pml_widths=pml,
                     use_reduced_precision=reduced, domain_zz=Z)
  ports = [("x", 6, True), ("x", xx - 8, True)]
  return eps, ports, params, omega


def coupler(total=(384, 256, 128), pad=32, pml=(16, 16), width=12, thick=8, gap=6, tt=20000,
            dt=0.5, reduced=False):
  """cfg4: four parallel waveguides, 8 x-plane ports (both ends of each), one run per port."""
  X, Y, Z = total
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  zc = zz / 2
  pitch = width + gap
  centres = yy / 2 + (np.arange(4) - 1.5) * pitch * 2
  centres_mid = yy / 2 + (np.arange(4) - 1.5) * pitch   # arms approach each other mid-way

  def core(x, y, z):
    tx = np.clip((x - xx * 0.25) / (xx * 0.15), 0, 1) - np.clip((x - xx * 0.6) / (xx * 0.15), 0, 1)
    s = 0.5 - 0.5 * np.cos(np.pi * tx)
    hit = False
    for c0, c1 in zip(centres, centres_mid):
      hit = hit | (np.abs(y - (c0 + (c1 - c0) * s)) < width / 2)
    return hit & (np.abs(z - zc) < thick / 2)

  eps = _yee(core, xx, yy, zz)
  params = SimParams(omega_range=(OMEGA0, OMEGA0), tt=tt, dt=dt, absorption_padding=pad,
                     absorption_coeff=1e-4, pml_widths=pml, use_reduced_precision=reduced,
                     domain_zz=Z)
  ports = [("x", 6, True)] * 4 + [("x", xx - 8, False)] * 4
  return eps, ports, params, np.array([OMEGA0]), centres


def metalens(total=(4096, 4096, 128), pad=32, pml=(16, 16), pitch=16, height=24, tt=2000,
             dt=0.5, seed=1, reduced=False):
  """cfg5: pillar-lattice metalens on a substrate, z-plane plane-wave source."""
  X, Y, Z = total
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  rng = np.random.default_rng(seed)
  radii = rng.uniform(2, 7, (xx // pitch + 1, yy // pitch + 1)).astype(np.float32)
  zsub = zz // 3

  def core(x, y, z):
    ix = np.floor(x / pitch).astype(np.int64).clip(0, radii.shape[0] - 1)
    iy = np.floor(y / pitch).astype(np.int64).clip(0, radii.shape[1] - 1)
    cx, cy = (ix + 0.5) * pitch, (iy + 0.5) * pitch
    r = radii[ix, iy]
    pillar = ((x - cx) ** 2 + (y - cy) ** 2 < r ** 2) & (z >= zsub) & (z < zsub + height)
    return pillar

  eps = _yee(core, xx, yy, zz)
  eps[:, :, :, :zsub] = np.where(eps[:, :, :, :zsub] == EPS_SI, EPS_SI, EPS_CLAD)
  eps[:, :, :, zsub + height:] = 1.0
  eps[:, :, :, zsub:zsub + height] = np.where(eps[:, :, :, zsub:zsub + height] == EPS_SI,
                                              EPS_SI, 1.0)
  params = SimParams(omega_range=(OMEGA0, OMEGA0), tt=tt, dt=dt, absorption_padding=pad,
                     absorption_coeff=1e-4, pml_widths=pml, use_reduced_precision=reduced,
                     domain_zz=Z)
  ports = [("z", zsub // 2, True)]
  return eps, ports, params, np.array([OMEGA0])


def metalens_columns(total, ycols, device, pad=32, pml=(16, 16), pitch=16, height=24, seed=1):
  """The permittivity of ``metalens(total, ...)`` restricted to the user-domain columns ``ycols``
  (clipped indices, i.e. edge replication already applied), built with torch on ``device``: one
  rank of the decomposed cfg5 run builds only ITS slab -- the 4096x4096x128 stack (18 GB of
  float32) is never materialised.  Same pillar lattice, radii and layers as ``metalens``."""
  import torch
  X, Y, Z = total
  xx, yy, zz = X - 2 * pad, Y - 2 * pad, Z - sum(pml)
  rng = np.random.default_rng(seed)
  radii = torch.from_numpy(
      rng.uniform(2, 7, (xx // pitch + 1, yy // pitch + 1)).astype(np.float32)).to(device)
  zsub = zz // 3
  yc = torch.as_tensor(np.asarray(ycols), dtype=torch.float32, device=device)
  xs = torch.arange(xx, dtype=torch.float32, device=device)
  eps = torch.empty((3, xx, yc.numel(), zz), dtype=torch.float32, device=device)
  shifts = ((0.5, 0, 0), (0, 0.5, 0), (0, 0, 0.5))
  for c, (sx, sy, sz) in enumerate(shifts):
    x, y = (xs + sx)[:, None], (yc + sy)[None, :]
    ix = torch.floor(x / pitch).long().clamp(0, radii.shape[0] - 1)
    iy = torch.floor(y / pitch).long().clamp(0, radii.shape[1] - 1)
    cx, cy = (ix + 0.5) * pitch, (iy + 0.5) * pitch
    inside = ((x - cx) ** 2 + (y - cy) ** 2 < radii[ix, iy] ** 2)          # (xx, ny)
    z = torch.arange(zz, dtype=torch.float32, device=device) + sz
    layer = ((z >= zsub) & (z < zsub + height))[None, None, :]
    pillar = inside[:, :, None] & layer
    # metalens(): substrate below zsub, pillars or air in the layer, air above -- indexed by the
    # INTEGER z plane (the reference builder overwrites whole planes after sampling)
    zi = torch.arange(zz, device=device)
    below = (zi < zsub)[None, None, :]
    in_layer = ((zi >= zsub) & (zi < zsub + height))[None, None, :]
    val = torch.where(below, torch.tensor(EPS_CLAD, device=device),
                      torch.where(in_layer & pillar, torch.tensor(EPS_SI, device=device),
                                  torch.tensor(1.0, device=device)))
    eps[c] = val
  return eps


def port_mode(eps, axis, pos, omega, num_modes=1):
  """Mode of the cross-section of ``eps`` at the port plane (harness: pjz_b200.mode)."""
  from ._mode import mode
  a = "xyz".index(axis)
  sl = [slice(None)] * 4
  sl[a + 1] = slice(pos, pos + 1)
  return mode(eps[tuple(sl)], np.atleast_1d(omega), num_modes)


def gaussian_port_source(eps, axis, pos, width=8.0):
  """Cheap stand-in for a mode profile (benchmarks that must not depend on ARPACK timing):
  a Gaussian beam profile on the port plane, transverse component 0 only."""
  a = "xyz".index(axis)
  shape = list(eps.shape[1:])
  shape[a] = 1
  idx = [np.arange(n, dtype=np.float32) for n in eps.shape[1:]]
  grids = np.meshgrid(*idx, indexing="ij", sparse=True)
  r2 = 0
  for i in range(3):
    if i != a:
      sl = [slice(None)] * 4
      sl[a + 1] = slice(pos, pos + 1)
      plane = eps[tuple(sl)][0]
      w = (plane > (EPS_SI + EPS_CLAD) / 2).astype(np.float32)
      tot = max(float(w.sum()), 1.0)
      centre = float((w * grids[i]).sum() / tot) if w.sum() > 0 else eps.shape[i + 1] / 2
      r2 = r2 + ((grids[i] - centre) / width) ** 2
  prof = np.exp(-r2).astype(np.float32)
  prof = np.broadcast_to(prof, shape)
  src = np.zeros((2,) + tuple(shape), np.float32)
  src[0] = prof
  return src
