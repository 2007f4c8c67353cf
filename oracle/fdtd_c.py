"""ctypes front-end of the C fp32 oracle (``oracle/fdtd_c.c``).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see ``oracle/__init__.py``).  Same keyword signature as
``fdtdz_jax.fdtdz`` (/root/reference/src/pjz/_field.py:254-269).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .fdtd_numpy import domain_shape, source_axis

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_fdtd.so")


class _Desc(ctypes.Structure):
  _fields_ = [(n, ctypes.c_int32) for n in (
      "X", "Y", "Z", "xx", "yy", "zz", "ox", "oy", "oz", "tt", "src_axis", "src_pos",
      "pml_lo", "pml_hi", "out_start", "out_stop", "out_step", "reduced")] + [
          ("dt", ctypes.c_float)]


def build(force=False):
  src = os.path.join(_HERE, "fdtd_c.c")
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"],
                          stdout=subprocess.DEVNULL)
  return _SO


_lib = None


def lib():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build())
    _lib.oracle_fdtd_run.restype = ctypes.c_int
    _lib.oracle_fdtd_max_threads.restype = ctypes.c_int
  return _lib


def _f32(a):
  return np.ascontiguousarray(np.asarray(a, np.float32))


def _ptr(a):
  return a.ctypes.data_as(ctypes.c_void_p)


def max_threads():
  return int(lib().oracle_fdtd_max_threads())


def fdtdz(epsilon, dt, source_field, source_waveform, source_position, absorption_mask,
          pml_kappa, pml_sigma, pml_alpha, pml_widths, output_steps,
          use_reduced_precision=False, launch_params=None, offset=(0, 0, 0),
          nthreads=0, steps_override=-1, want_output=True, output_projection=None):
  eps, sf, wf = _f32(epsilon), _f32(source_field), _f32(source_waveform)
  mask, kap, sig, alp = (_f32(a) for a in (absorption_mask, pml_kappa, pml_sigma, pml_alpha))
  X, Y, Z = domain_shape(mask, kap)
  axis = source_axis(sf)
  expect = {0: (2, 1, Y, Z), 1: (2, X, 1, Z), 2: (2, 2, X, Y, 1)}[axis]
  if tuple(sf.shape) != expect:
    raise ValueError(f"source_field shape {sf.shape} != {expect}")
  d = _Desc()
  d.X, d.Y, d.Z = X, Y, Z
  _, d.xx, d.yy, d.zz = eps.shape
  d.ox, d.oy, d.oz = (int(o) for o in offset)
  d.tt = wf.shape[0]
  d.src_axis, d.src_pos = axis, int(source_position)
  d.pml_lo, d.pml_hi = int(pml_widths[0]), int(pml_widths[1])
  d.out_start, d.out_stop, d.out_step = (int(v) for v in output_steps)
  d.reduced = int(bool(use_reduced_precision))
  d.dt = float(dt)
  nout = len(range(*output_steps))
  out = np.zeros((nout, 3, d.xx, d.yy, d.zz), np.float32) if want_output else None
  rc = lib().oracle_fdtd_run(ctypes.byref(d), _ptr(eps), _ptr(sf), _ptr(wf), _ptr(mask),
                             _ptr(kap), _ptr(sig), _ptr(alp),
                             _ptr(out) if out is not None else None,
                             ctypes.c_int(nthreads), ctypes.c_int(steps_override))
  if rc != 0:
    raise ValueError(f"oracle_fdtd_run failed with code {rc}")
  if output_projection is not None and out is not None:
    w = _f32(output_projection)
    if w.ndim != 2 or w.shape[1] != nout:
      raise ValueError(f"output_projection must have shape (rows, {nout}), got {w.shape}")
    proj = np.zeros((w.shape[0],) + out.shape[1:], np.float32)
    n = int(np.prod(out.shape[1:]))
    lib().oracle_fdtd_project(_ptr(out), _ptr(w), ctypes.c_int(w.shape[0]), ctypes.c_int(nout),
                              ctypes.c_size_t(n), _ptr(proj))
    return proj
  return out
