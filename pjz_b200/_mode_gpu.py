"""Waveguide mode solver on the GPU.

Same algorithm, interface and return layout as ``pjz.mode``
(/root/reference/src/pjz/_mode.py:148-257): shifted subspace iteration (:101-145) on the
waveguide operator (:22-51), batched over the ``ww`` frequencies.  The operator is one
hand-written CUDA launch per application (``b200fdtd_mode_operator``, csrc/postproc.cuh); the
thin QR and residual norms are batched torch/cuSOLVER calls on the same stream.  Two changes
to the iteration itself, same fixed point: a Chebyshev polynomial filter replaces the plain
power step and a Rayleigh-Ritz projection (guard vectors included) replaces the Rayleigh
quotients, because the reference's loop converges at the ratio of adjacent shifted eigenvalues
(0.9993 on the golden waveguide) and in practice runs into ``max_iters``.  A warm start with a
converged ``init`` returns ``iters == 1`` like the reference
(/root/reference/tests/test_modes.py:50-73).

``pjz_b200.mode`` (``_mode.py``, ARPACK on the host) stays the harness the engine tests use;
this module is the device-resident variant for optimisation loops, where every step re-solves
every port.  Golden betas: /root/reference/tests/test_modes.py:34-47.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch


def _apply(eps, omega, shift, x):
  """y = op(x) through the CUDA kernel; x is (ww, 2, uu, vv, mm) float32, contiguous."""
  from . import fdtdz_jax as shim
  L = shim.lib()
  ww, _, uu, vv, mm = x.shape
  y = torch.empty_like(x)
  L.b200fdtd_mode_operator.restype = ctypes.c_int
  L.b200fdtd_mode_operator.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 6
  rc = L.b200fdtd_mode_operator(ww, uu, vv, mm, eps.data_ptr(), omega.data_ptr(),
                                shift.data_ptr(), x.data_ptr(), y.data_ptr(),
                                torch.cuda.current_stream(x.device).cuda_stream)
  if rc != 0:
    raise RuntimeError(f"b200fdtd_mode_operator failed ({rc}): {shim._last_error()}")
  return y


def _power_iteration(eps, omega, x, n):
  """Most negative eigenvalue of the unshifted operator (what the reference uses as ``shift``,
  :225-231): ``n`` normalised power steps.  Returns a lower bound estimate per frequency: the
  Rayleigh quotient minus twice the residual norm minus 10 % slack."""
  zeros = torch.zeros_like(omega)
  ww = x.shape[0]
  w = rho = None
  for _ in range(n):
    x = x / torch.linalg.norm(x.reshape(ww, -1), dim=1)[:, None, None, None, None]
    y = _apply(eps, omega, zeros, x)
    w = torch.sum(x * y, dim=(1, 2, 3, 4))
    rho = torch.linalg.norm((y - w[:, None, None, None, None] * x).reshape(ww, -1), dim=1)
    x = y
  return w - 2.0 * rho - 0.1 * w.abs()


_HOST_EIG = False     # set when this torch build has no batched CUDA geev (see _ritz)


def _ritz(q, aq):
  """Rayleigh-Ritz on the block ``q`` (ww, N, p) with ``aq = op(q)``.  The waveguide operator is
  not symmetric (the permittivity enters on one side only), so the small projected matrices
  (p <= ~16) go through a general eigen-solver: ``torch.linalg.eig`` batched over the frequencies
  ON THE DEVICE (as the reference keeps its Rayleigh quotients inside ``lax.while_loop``,
  /root/reference/src/pjz/_mode.py:129-138) -- nothing is copied to the host.  Eigenvalues are real
  for lossless media.  Returns Ritz values (descending), unit-norm Ritz vectors, op(Ritz vectors)."""
  global _HOST_EIG
  t = torch.matmul(q.transpose(1, 2), aq).double()
  lam = vec = None
  if not _HOST_EIG:
    try:
      lam, vec = torch.linalg.eig(t)
    except RuntimeError:                                   # no device geev in this build
      _HOST_EIG = True
  if lam is None:
    lam, vec = torch.linalg.eig(t.cpu())
    lam, vec = lam.to(q.device), vec.to(q.device)
  lam_r, vec_r = lam.real, vec.real
  order = torch.argsort(lam_r, dim=1, descending=True)
  theta = torch.gather(lam_r, 1, order).to(q.dtype)
  s = torch.gather(vec_r, 2, order[:, None, :].expand_as(vec_r)).to(q.dtype)
  x, ax = torch.matmul(q, s), torch.matmul(aq, s)
  nrm = torch.linalg.norm(x, dim=1, keepdim=True)
  return theta, x / nrm, ax / nrm


def _subspace_iteration(eps, omega, lam_min, x, n, tol, keep, degree=40):
  """Chebyshev-filtered subspace iteration for the ``keep`` LARGEST eigenpairs.

  The reference (:101-145) iterates ``x <- orth(op(x))`` on the operator shifted by its most
  negative eigenvalue; each vector then converges at the ratio of adjacent shifted eigenvalues,
  0.9993 for the golden waveguide, and the loop simply runs into ``max_iters``.  Same fixed
  point, far fewer operator applications: every outer iteration applies a degree-``degree``
  Chebyshev polynomial of the operator that is bounded on ``[lam_min, cut]`` (everything below
  the current block) and grows exponentially above it, then a Rayleigh-Ritz step.  ``x`` may
  carry guard vectors beyond ``keep``.  Returns ``(w, x, err, outer iterations)``."""
  ww, p = x.shape[0], x.shape[-1]
  shape = x.shape
  zeros = torch.zeros_like(omega)
  op = lambda v, c: _apply(eps, omega, c, v.contiguous())
  q, _ = torch.linalg.qr(x.reshape(ww, -1, p), mode="reduced")
  theta, xr, axr = _ritz(q, op(q.reshape(shape), zeros).reshape(ww, -1, p))
  err = torch.linalg.norm(axr - theta[:, None, :] * xr, dim=1)
  lo = lam_min.clone()
  i = 0
  while i < n:
    i += 1
    if i > 1 or float(err[:, :keep].max()) > tol:
      for _attempt in range(6):
        cut = theta[:, -1]                                   # damp everything below the block
        c = (0.5 * (cut + lo)).contiguous()
        e = 0.5 * (cut - lo)
        # scaled three-term recurrence (normalised at the top Ritz value: no overflow in fp32)
        b5 = lambda t: t[:, None, None, None, None]
        sig1 = e / torch.clamp(theta[:, 0] - c, min=1e-6)
        sig = sig1
        y0 = xr.reshape(shape)
        y1 = op(y0, c) * b5(sig1 / e)
        for _ in range(2, degree + 1):
          sig2 = 1.0 / (2.0 / sig1 - sig)
          y0, y1 = y1, op(y1, c) * b5(2.0 * sig2 / e) - y0 * b5(sig * sig2)
          sig = sig2
        if bool(torch.isfinite(y1).all()):
          break
        lo = lo - 0.5 * lo.abs()                             # spectrum reaches below the bound: widen
      q, _ = torch.linalg.qr(y1.reshape(ww, -1, p), mode="reduced")
      theta, xr, axr = _ritz(q, op(q.reshape(shape), zeros).reshape(ww, -1, p))
      err = torch.linalg.norm(axr - theta[:, None, :] * xr, dim=1)
    if float(err[:, :keep].max()) <= tol:
      break
  return theta[:, :keep], xr.reshape(shape)[..., :keep].contiguous(), err[:, :keep], i


def _diff(arr, axis, is_forward):
  if is_forward:
    return torch.roll(arr, -1, axis) - arr
  return arr - torch.roll(arr, 1, axis)


def _poynting(beta, omega, eps, x):
  """z-directed power of every mode (:73-87); x is (ww, 2, uu, vv, mm), eps (3, uu, vv)."""
  b = beta[:, None, None, :]
  om = omega[:, None, None, None]
  hx, hy = x[:, 0].to(torch.complex64), x[:, 1].to(torch.complex64)
  hz = (_diff(x[:, 0], -3, True) + _diff(x[:, 1], -2, True)) / (1j * b)
  dz = lambda f: -1j * b * f
  # e = curl_backward(h) / (i omega eps): only the transverse pair enters the power
  ex = (_diff(hz, -2, False) - dz(hy)) / (1j * om * eps[0][None, :, :, None])
  ey = (dz(hx) - _diff(hz, -3, False)) / (1j * om * eps[1][None, :, :, None])
  return torch.real(torch.sum(ex * hy - ey * hx, dim=(1, 2)))


def mode_gpu(epsilon, omega, num_modes: int, init: Optional[torch.Tensor] = None,
             shift_iters: int = 10, max_iters: int = 100000, tol: float = 1e-4,
             guard: Optional[int] = None, device=None):
  """Solve for waveguide modes on the GPU; interface of ``pjz.mode`` (:148-196).

  ``guard`` extra trial vectors (default ``max(4, num_modes)``) ride along in the iterated block
  and are dropped at the end.  Returns ``(wavevector (ww, mm), excitation (ww, 2, xx, yy, zz, mm),
  err (ww, mm), iters)`` as float32 CUDA tensors (``iters`` a python int); index 0 is the
  fundamental mode."""
  if not torch.cuda.is_available():
    raise RuntimeError("mode_gpu needs a CUDA device (no CPU fallback)")
  dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
  eps = torch.as_tensor(np.asarray(epsilon) if not isinstance(epsilon, torch.Tensor) else epsilon)
  eps = eps.to(dev, torch.float32)
  omega = torch.atleast_1d(torch.as_tensor(np.asarray(omega) if not isinstance(omega, torch.Tensor)
                                           else omega)).to(dev, torch.float32)
  if 1 not in tuple(eps.shape[1:]):
    raise ValueError(
        f"Expected exactly one of the spatial dimensions of ``epsilon`` to be "
        f"singular, instead got ``epsilon.shape == {tuple(eps.shape)}``.")
  prop_axis = "xyz"[tuple(eps.shape).index(1, 1) - 1]
  ww = omega.shape[0]

  # "Propagate-along-z" form (:211-217).
  if prop_axis == "x":
    e2 = eps[[1, 2, 0]]
  elif prop_axis == "y":
    e2 = torch.flip(torch.swapaxes(eps[[2, 0, 1]], 1, 3), dims=(1,))
  else:
    e2 = eps
  e2 = torch.squeeze(e2, dim=tuple(i for i in (1, 2, 3) if e2.shape[i] == 1)).contiguous()
  _, uu, vv = e2.shape
  mode_shape = (ww, 2, uu, vv, num_modes)
  guard = max(4, num_modes) if guard is None else int(guard)
  guard = max(0, min(guard, 2 * uu * vv - num_modes))

  gen = torch.Generator(device="cpu").manual_seed(0)
  if init is None:
    x0 = torch.randn(mode_shape, generator=gen).to(dev)
  else:
    x0 = torch.as_tensor(init).to(dev, torch.float32)
    x0 = torch.squeeze(x0, dim="xyz".index(prop_axis) + 2)
    if prop_axis == "y":                                   # output form -> solver form (:199-206)
      x0 = torch.flip(torch.swapaxes(x0, 2, 3), dims=(1, 2))
    elif prop_axis == "z":
      x0 = x0 * torch.tensor([1.0, -1.0], device=dev)[None, :, None, None, None]
    x0 = torch.flip(x0, dims=(1,)).reshape(mode_shape).contiguous()
  if guard:
    x0 = torch.cat([x0, torch.randn(mode_shape[:-1] + (guard,), generator=gen).to(dev)], dim=-1)

  with torch.cuda.device(dev):
    lam_min = _power_iteration(e2, omega, torch.randn(mode_shape[:-1] + (1,), generator=gen).to(dev),
                               max(shift_iters, 40))
    w, x, err, iters = _subspace_iteration(e2, omega, lam_min, x0.contiguous(), max_iters, tol,
                                           keep=num_modes)
    beta = torch.sqrt(torch.clamp(w, min=0))
    p = _poynting(beta, omega, e2, x)
    x = x / torch.sqrt(torch.abs(p))[:, None, None, None, :]

  exc = torch.flip(x, dims=(1,))                            # field -> excitation (:243)
  if prop_axis == "y":
    exc = torch.swapaxes(torch.flip(exc, dims=(1, 2)), 2, 3)
  elif prop_axis == "z":
    exc = exc * torch.tensor([1.0, -1.0], device=dev)[None, :, None, None, None]
  exc = torch.unsqueeze(exc, "xyz".index(prop_axis) + 2)
  return beta.to(torch.float32), exc.to(torch.float32).contiguous(), err.to(torch.float32), iters
