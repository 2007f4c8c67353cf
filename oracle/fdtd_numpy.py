"""NumPy statement of the engine semantics (the executable spec).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: see ``oracle/__init__.py``.  This file *defines* what
``fdtdz_jax.fdtdz`` (call site /root/reference/src/pjz/_field.py:254-269) computes for
this repo; SURVEY.md section 8(c) is the spec it freezes.  Conventions taken from the
reference tree:

* Yee staggering: Ex,Ey (and Hz) on integer-z planes, Ez (and Hx,Hy) on half-z planes
  (/root/reference/src/pjz/_epsilon.py:13-24,151-152).
* curl E uses forward differences, curl H backward differences, x-y periodic
  (/root/reference/src/pjz/_mode.py:15-19,54-98).
* ``pml_*`` are (Z,2): column 0 at integer z (E-type), column 1 at z+1/2 (H-type)
  (/root/reference/src/pjz/_field.py:98-106, /root/reference/tests/test_boundaries.py:27-41).
* x/y plane sources use waveform channel 0 on plane ``source_position`` and channel 1 on
  plane ``source_position-1`` (deduced from the odd-position fix-up,
  /root/reference/src/pjz/_field.py:230-233); z sources are (2,2,X,Y,1) = [channel][Ex,Ey]
  (/root/reference/src/pjz/_field.py:234-236).
* the source is added to E with no 1/epsilon factor (pjz pre-divides,
  /root/reference/src/pjz/_field.py:166-167,219).
* output = E after update+source at steps ``range(*output_steps)``, cropped to epsilon's
  sub-volume at ``offset`` (/root/reference/src/pjz/_field.py:268,276-279).

Units dx=dy=dz=1 (dz stretched by kappa), mu=1, c=1.
"""

from __future__ import annotations

import numpy as np


def domain_shape(absorption_mask, pml_kappa):
  """Full simulation domain (X, Y, Z): x-y from the absorber mask, z from the PML tables."""
  return (int(absorption_mask.shape[1]), int(absorption_mask.shape[2]),
          int(pml_kappa.shape[0]))


def source_axis(source_field):
  """0/1/2 for x/y/z plane sources, from the source array's shape."""
  shp = tuple(source_field.shape)
  if len(shp) == 5:
    if shp[0] != 2 or shp[1] != 2 or shp[4] != 1:
      raise ValueError(f"z source must be (2,2,X,Y,1), got {shp}")
    return 2
  if len(shp) == 4 and shp[0] == 2:
    if shp[1] == 1:
      return 0
    if shp[2] == 1:
      return 1
  raise ValueError(
      f"source_field must be (2,1,Y,Z), (2,X,1,Z) or (2,2,X,Y,1), got {shp}")


def cpml_tables(pml_kappa, pml_sigma, pml_alpha, pml_widths, dt):
  """Per-z CPML coefficients, float64: dict of (Z,) arrays a_e,b_e,ik_e,a_h,b_h,ik_h.

  b = exp(-(sigma/kappa + alpha) dt);  a = sigma (b-1) / (kappa (sigma + kappa alpha)),
  a := 0 where sigma == 0 and outside the ``pml_widths`` cells;  1/kappa := 0 for
  kappa = inf (``use_z_as_batch``, /root/reference/src/pjz/_field.py:243-246).
  """
  kappa = np.asarray(pml_kappa, np.float64)
  sigma = np.asarray(pml_sigma, np.float64)
  alpha = np.asarray(pml_alpha, np.float64)
  zz = kappa.shape[0]
  z = np.arange(zz)
  in_pml = (z < pml_widths[0]) | (z >= zz - pml_widths[1])
  out = {}
  for col, tag in ((0, "e"), (1, "h")):
    k, s, al = kappa[:, col], sigma[:, col], alpha[:, col]
    with np.errstate(divide="ignore", invalid="ignore"):
      ik = np.where(np.isinf(k), 0.0, 1.0 / k)
      b = np.exp(-(s * ik + al) * dt)
      denom = k * (s + k * al)
      a = np.where((s != 0) & in_pml & np.isfinite(denom) & (denom != 0),
                   s * (b - 1.0) / np.where(denom == 0, 1.0, denom), 0.0)
    a = np.where(np.isfinite(a), a, 0.0)
    out["a_" + tag], out["b_" + tag], out["ik_" + tag] = a, b, ik
  return out


def absorber_coeffs(absorption_mask, dt):
  """(A, S) float64 (3,X,Y): A=(1-s dt/2)/(1+s dt/2), S=1/(1+s dt/2)."""
  s = np.asarray(absorption_mask, np.float64)
  return (1 - s * dt / 2) / (1 + s * dt / 2), 1 / (1 + s * dt / 2)


def extend_epsilon(epsilon, full_shape, offset):
  """Edge-replicate the (3,xx,yy,zz) sub-volume to the full (3,X,Y,Z) domain."""
  eps = np.asarray(epsilon)
  pads = [(0, 0)]
  for i in range(3):
    lo = int(offset[i])
    hi = int(full_shape[i]) - lo - eps.shape[i + 1]
    if lo < 0 or hi < 0:
      raise ValueError("epsilon sub-volume at offset does not fit the domain")
    pads.append((lo, hi))
  return np.pad(eps, pads, mode="edge")


def _dfwd(a, axis):
  return np.roll(a, -1, axis) - a


def _dbwd(a, axis):
  return a - np.roll(a, 1, axis)


def _dz_fwd(a):
  """a[z+1]-a[z] with a[Z] := 0."""
  out = -a.copy()
  out[..., :-1] += a[..., 1:]
  return out


def _dz_bwd(a):
  """a[z]-a[z-1] with a[-1] := 0."""
  out = a.copy()
  out[..., 1:] -= a[..., :-1]
  return out


class State:
  """Mutable simulation state + precomputed coefficients."""

  def __init__(self, epsilon, dt, absorption_mask, pml_kappa, pml_sigma, pml_alpha,
               pml_widths, offset=(0, 0, 0), dtype=np.float64, storage=None):
    self.dtype = np.dtype(dtype)
    self.storage = storage  # e.g. np.float16: fields/coefficient rounded to it when stored
    self.dt = float(np.float32(dt))
    self.shape = domain_shape(absorption_mask, pml_kappa)
    X, Y, Z = self.shape
    eps = extend_epsilon(np.asarray(epsilon, np.float32), self.shape, offset)
    A, S = absorber_coeffs(np.asarray(absorption_mask, np.float32), self.dt)
    self.A = A.astype(dtype)[..., None]                      # (3,X,Y,1)
    S = S.astype(np.float32).astype(dtype)[..., None]
    dt_t = self.dtype.type(np.float32(self.dt))
    self.B = ((dt_t / eps.astype(dtype)) * S).astype(dtype)  # (3,X,Y,Z)
    if storage is not None:
      self.B = self.B.astype(storage).astype(dtype)
    t = cpml_tables(np.asarray(pml_kappa, np.float32), np.asarray(pml_sigma, np.float32),
                    np.asarray(pml_alpha, np.float32), pml_widths, self.dt)
    # Tables are float32-rounded first (the engine holds them as float32).
    self.t = {k: v.astype(np.float32).astype(dtype) for k, v in t.items()}
    self.E = np.zeros((3, X, Y, Z), dtype)
    self.H = np.zeros((3, X, Y, Z), dtype)
    self.psiH = np.zeros((2, X, Y, Z), dtype)   # psiHx, psiHy
    self.psiE = np.zeros((2, X, Y, Z), dtype)   # psiEx, psiEy
    self.dt_t = dt_t

  def _store(self, a):
    if self.storage is None:
      return a
    return a.astype(self.storage).astype(self.dtype)

  def step_h(self):
    E, H, t, dt = self.E, self.H, self.t, self.dt_t
    dzEy, dzEx = _dz_fwd(E[1]), _dz_fwd(E[0])
    self.psiH[0] = t["b_h"] * self.psiH[0] + t["a_h"] * dzEy
    self.psiH[1] = t["b_h"] * self.psiH[1] + t["a_h"] * dzEx
    cx = _dfwd(E[2], 1) - (dzEy * t["ik_h"] + self.psiH[0])
    cy = (dzEx * t["ik_h"] + self.psiH[1]) - _dfwd(E[2], 0)
    cz = _dfwd(E[1], 0) - _dfwd(E[0], 1)
    H[0] = self._store(H[0] - dt * cx)
    H[1] = self._store(H[1] - dt * cy)
    H[2] = self._store(H[2] - dt * cz)

  def step_e(self):
    E, H, t = self.E, self.H, self.t
    dzHy, dzHx = _dz_bwd(H[1]), _dz_bwd(H[0])
    self.psiE[0] = t["b_e"] * self.psiE[0] + t["a_e"] * dzHy
    self.psiE[1] = t["b_e"] * self.psiE[1] + t["a_e"] * dzHx
    cx = _dbwd(H[2], 1) - (dzHy * t["ik_e"] + self.psiE[0])
    cy = (dzHx * t["ik_e"] + self.psiE[1]) - _dbwd(H[2], 0)
    cz = _dbwd(H[1], 0) - _dbwd(H[0], 1)
    for c, cc in enumerate((cx, cy, cz)):
      E[c] = self.A[c] * E[c] + self.B[c] * cc

  def add_source(self, source_field, wf_row, source_position, axis):
    """E_transverse += sum_ch wf[ch] * source_field[ch] on the source plane(s)."""
    E = self.E
    sf = source_field
    p = int(source_position)
    X, Y, Z = self.shape
    if axis == 0:
      for ch in range(2):
        xp = (p - ch) % X
        E[1][xp] += wf_row[ch] * sf[0, 0]
        E[2][xp] += wf_row[ch] * sf[1, 0]
    elif axis == 1:
      for ch in range(2):
        yp = (p - ch) % Y
        E[0][:, yp] += wf_row[ch] * sf[0, :, 0]
        E[2][:, yp] += wf_row[ch] * sf[1, :, 0]
    else:
      for ch in range(2):
        E[0][:, :, p] += wf_row[ch] * sf[ch, 0, :, :, 0]
        E[1][:, :, p] += wf_row[ch] * sf[ch, 1, :, :, 0]

  def finish_e(self):
    if self.storage is not None:
      self.E[...] = self._store(self.E)


def fdtdz(epsilon, dt, source_field, source_waveform, source_position, absorption_mask,
          pml_kappa, pml_sigma, pml_alpha, pml_widths, output_steps,
          use_reduced_precision=False, launch_params=None, offset=(0, 0, 0),
          dtype=np.float64, return_state=False, output_projection=None):
  """Oracle with the ``fdtdz_jax.fdtdz`` keyword signature (/root/reference/src/pjz/_field.py:254-269).

  Returns float32 ``(n_out, 3, xx, yy, zz)``.  ``dtype`` selects the arithmetic precision of
  the oracle itself (float64 = truth, float32 = same-precision comparison);
  ``use_reduced_precision`` rounds stored E, H and the dt/epsilon coefficient to float16.
  """
  epsilon = np.asarray(epsilon, np.float32)
  source_field = np.asarray(source_field, np.float32).astype(dtype)
  source_waveform = np.asarray(source_waveform, np.float32).astype(dtype)
  axis = source_axis(source_field)
  st = State(epsilon, dt, absorption_mask, pml_kappa, pml_sigma, pml_alpha, pml_widths,
             offset, dtype, storage=np.float16 if use_reduced_precision else None)
  X, Y, Z = st.shape
  expect = {0: (2, 1, Y, Z), 1: (2, X, 1, Z), 2: (2, 2, X, Y, 1)}[axis]
  if tuple(source_field.shape) != expect:
    raise ValueError(f"source_field shape {source_field.shape} != {expect}")
  tt = source_waveform.shape[0]
  outs = list(range(*output_steps))
  if outs and (outs[0] < 0 or outs[-1] >= tt):
    raise ValueError("output_steps outside [0, tt)")
  _, xx, yy, zz = epsilon.shape
  ox, oy, oz = (int(o) for o in offset)
  out = np.zeros((len(outs), 3, xx, yy, zz), np.float32)
  oi = 0
  for n in range(tt):
    st.step_h()
    st.step_e()
    st.add_source(source_field, source_waveform[n], source_position, axis)
    st.finish_e()
    if oi < len(outs) and n == outs[oi]:
      out[oi] = st.E[:, ox:ox + xx, oy:oy + yy, oz:oz + zz]
      oi += 1
  if output_projection is not None:
    # engine extension: out[r] = sum_s W[r, s] * snapshot_s  (cf. _field.py:276-277)
    w = np.asarray(output_projection, np.float32).astype(np.float64)
    out = np.einsum("rs,s...->r...", w, out.astype(np.float64)).astype(np.float32)
  if return_state:
    return out, st
  return out
