"""Waveguide mode solver -- host-side harness that defines the ports fed to the engine.

Mirrors ``pjz.mode`` (/root/reference/src/pjz/_mode.py:148-257): same operator
(:22-51), same Poynting normalisation (:84-87), same field->excitation flips and axis
permutations (:199-253), same return layout ``(wavevector, excitation, err, iters)``.
The reference finds the top eigenpairs by shifted subspace iteration inside JAX
(:101-145); that loop is not on the engine hot path, so here the
same eigenproblem is handed to ARPACK (``scipy.sparse.linalg.eigs``, largest real part).
Golden betas: /root/reference/tests/test_modes.py:34-47.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.sparse.linalg import LinearOperator, eigs


def _diff(arr, axis, is_forward):
  """Periodic forward/backward difference (/root/reference/src/pjz/_mode.py:15-19)."""
  if is_forward:
    return np.roll(arr, -1, axis) - arr
  return arr - np.roll(arr, 1, axis)


def _apply_operator(epsilon, omega, arr):
  """Waveguide operator on ``arr`` (2, xx, yy) for one omega (:22-51, shift = 0)."""
  eps_yx = epsilon[(1, 0), :, :]
  eps_z = epsilon[2]
  a = (omega**2 * eps_yx) * arr
  b = -_diff(arr[0], -1, False) + _diff(arr[1], -2, False)
  b = b / eps_z
  b = np.stack([-_diff(b, -1, True), _diff(b, -2, True)], axis=0)
  b = b * eps_yx
  c = _diff(arr[0], -2, True) + _diff(arr[1], -1, True)
  c = np.stack([_diff(c, -2, False), _diff(c, -1, False)], axis=0)
  return a + b + c


def _fullh(beta, x):
  """(Hx, Hy, Hz) from the transverse pair (:73-81); x is (2, xx, yy)."""
  return np.stack([x[0], x[1],
                   (_diff(x[0], -2, True) + _diff(x[1], -1, True)) / (1j * beta)], axis=0)


def _curl(beta, arr, is_forward):
  """Curl with d/dz -> -i beta (:54-70); arr is (3, xx, yy)."""
  dx = lambda f: _diff(f, -2, is_forward)
  dy = lambda f: _diff(f, -1, is_forward)
  dz = lambda f: -1j * beta * f
  fx, fy, fz = arr
  return np.stack([dy(fz) - dz(fy), dz(fx) - dx(fz), dx(fy) - dy(fx)], axis=0)


def _full_fields(beta, omega, epsilon, x):
  """(h, e, h2) self-consistency triple (:90-98)."""
  h = _fullh(beta, x)
  e = _curl(beta, h, False) / (1j * omega * epsilon)
  h2 = _curl(beta, e, True) / (-1j * omega)
  return h, e, h2


def _poynting(beta, omega, epsilon, x):
  """z-directed power of the mode (:84-87)."""
  h = _fullh(beta, x)
  e = _curl(beta, h, False) / (1j * omega * epsilon)
  return np.real(np.sum(e[0] * h[1] - e[1] * h[0]))


def mode(epsilon, omega, num_modes: int, init: Optional[np.ndarray] = None,
         shift_iters: int = 10, max_iters: int = 100000, tol: float = 1e-4):
  """Solve for waveguide modes; interface of ``pjz.mode`` (:148-196).

  Args:
    epsilon: ``(3, xx, yy, zz)`` with exactly one singleton spatial dimension.
    omega: ``(ww,)`` angular frequencies.
    num_modes: number of modes.
    init, shift_iters, max_iters: accepted for signature parity; ARPACK needs no warm start.
    tol: eigen-solver tolerance.

  Returns:
    ``(wavevector (ww, num_modes), excitation (ww, 2, xx, yy, zz, num_modes) float32,
    err (ww, num_modes), iters)``; index 0 is the fundamental mode.
  """
  epsilon = np.asarray(epsilon, np.float64)
  omega = np.atleast_1d(np.asarray(omega, np.float64))
  if 1 not in epsilon.shape[1:]:
    raise ValueError(
        f"Expected exactly one of the spatial dimensions of ``epsilon`` to be "
        f"singular, instead got ``epsilon.shape == {epsilon.shape}``.")
  prop_axis = "xyz"[epsilon.shape.index(1, 1) - 1]

  # "Propagate-along-z" form (:211-217).
  if prop_axis == "x":
    eps = epsilon[(1, 2, 0), ...]
  elif prop_axis == "y":
    eps = np.flip(np.swapaxes(epsilon[(2, 0, 1), ...], 1, 3), axis=1)
  else:
    eps = epsilon
  eps = np.squeeze(eps, axis=tuple(i for i in (1, 2, 3) if eps.shape[i] == 1))
  _, uu, vv = eps.shape
  n = 2 * uu * vv

  betas, excs, errs = [], [], []
  for w in omega:
    op = LinearOperator(
        (n, n), dtype=np.float64,
        matvec=lambda v, w=w: _apply_operator(eps, w, v.reshape(2, uu, vv)).reshape(-1))
    rng = np.random.default_rng(0)
    vals, vecs = eigs(op, k=num_modes, which="LR", tol=min(tol, 1e-8) * 1e-2,
                      v0=rng.standard_normal(n), maxiter=max(10 * n, 20000))
    order = np.argsort(-vals.real)
    vals, vecs = vals.real[order], vecs.real[:, order]
    beta = np.sqrt(vals)
    xs, err = [], []
    for k in range(num_modes):
      x = vecs[:, k].reshape(2, uu, vv)
      x = x / np.linalg.norm(x)
      err.append(np.linalg.norm(_apply_operator(eps, w, x) - vals[k] * x))
      p = _poynting(beta[k], w, eps, x)
      x = x / np.sqrt(abs(p))
      if p < 0:
        x = x  # sign of the eigenvector is arbitrary; power sign handled by abs.
      xs.append(x)
    betas.append(beta)
    errs.append(err)
    excs.append(np.stack(xs, axis=-1))                   # (2, uu, vv, mm)
  x = np.stack(excs, axis=0)                             # (ww, 2, uu, vv, mm)

  exc = np.flip(x, axis=1)                               # field -> excitation (:243)
  if prop_axis == "y":
    exc = np.swapaxes(np.flip(exc, axis=(1, 2)), 2, 3)
  elif prop_axis == "z":
    exc = exc * np.array([1, -1])[None, :, None, None, None]
  exc = np.expand_dims(exc, "xyz".index(prop_axis) + 2)
  return (np.asarray(betas, np.float32), np.ascontiguousarray(exc, np.float32),
          np.asarray(errs, np.float32), 0)
