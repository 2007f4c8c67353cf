"""Seeded synthetic engine inputs shared by the tests (kwargs of ``fdtdz_jax.fdtdz``,
/root/reference/src/pjz/_field.py:254-269)."""

import numpy as np

from pjz_b200 import _field as glue


def random_problem(domain=(12, 10, 16), sub=None, offset=None, axis=0, pml=(3, 4), tt=12,
                   output_steps=None, seed=0, dt=0.5, z_as_batch=False, reduced=False,
                   src_pos=None, absorb_pad=3, absorb_coeff=2e-2):
  """The SURVEY 8(d) "stress" case generator: random epsilon in [1, 12.25], random source field,
  real absorber mask / PML sigma profiles from the pjz builders."""
  rng = np.random.default_rng(seed)
  X, Y, Z = domain
  if sub is None:
    sub = (max(1, X - 4), max(1, Y - 3), max(1, Z - 5))
  if offset is None:
    offset = tuple(int(rng.integers(0, d - s + 1)) for d, s in zip(domain, sub))
  eps = rng.uniform(1.0, 12.25, (3,) + tuple(sub)).astype(np.float32)
  shape = {0: (2, 1, Y, Z), 1: (2, X, 1, Z), 2: (2, 2, X, Y, 1)}[axis]
  sf = rng.standard_normal(shape).astype(np.float32)
  t = np.arange(tt)
  wf = np.stack([np.sin(0.31 * t) * (1 - np.exp(-t / 5.0)),
                 0.5 * np.cos(0.27 * t) * (1 - np.exp(-t / 4.0))], -1).astype(np.float32)
  mask = glue._absorption_mask(X, Y, absorb_pad, absorb_coeff)
  if z_as_batch:
    kappa = np.full((Z, 2), np.inf, np.float32)
    sigma = np.zeros((Z, 2), np.float32)
    alpha = np.zeros((Z, 2), np.float32)
  else:
    sigma = glue._pml_sigma(pml, Z, 0.5, 1.3)
    kappa = (1.0 + 0.3 * sigma).astype(np.float32)      # exercise kappa != 1 too
    alpha = np.full((Z, 2), 0.05, np.float32)
  ext = domain[axis]
  if src_pos is None:
    src_pos = int(rng.integers(0, ext))
  if output_steps is None:
    output_steps = (max(0, tt - 7), tt, 3)
  return dict(epsilon=eps, dt=dt, source_field=sf, source_waveform=wf,
              source_position=src_pos, absorption_mask=mask, pml_kappa=kappa, pml_sigma=sigma,
              pml_alpha=alpha, pml_widths=tuple(pml), output_steps=tuple(output_steps),
              use_reduced_precision=reduced, launch_params=None, offset=tuple(offset))


def rel_l2(a, b):
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  den = np.linalg.norm(b)
  return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a))
