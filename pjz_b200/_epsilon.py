"""Permittivity on the Yee cell, rendered on the GPU.

Mirrors ``pjz.render`` / ``pjz.epsilon`` (/root/reference/src/pjz/_epsilon.py:95-155): same
arguments, same ``(3, xx, yy, zz)`` result, produced by ``b200fdtd_render`` (csrc/render.cuh)
as a float32 CUDA tensor in the layout ``fdtdz_jax.fdtdz`` takes as ``epsilon`` -- no host
round trip between the renderer and the engine.  Differentiable (``torch.autograd``) with respect
to ``layers`` and ``layer_pos`` through ``b200fdtd_render_backward`` -- the reference obtains
these gradients with ``jax.grad`` (tests/test_layers.py:179-188) -- so that the chain
layers -> epsilon -> ``scatter`` -> loss runs backward on the GPU end to end.
Golden values: /root/reference/tests/test_layers.py.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch


def _dev_f32(a, dev):
  if isinstance(a, torch.Tensor):
    return a.to(dev, torch.float32).contiguous()
  return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).to(dev)


def _overlap_table(pos, gs, ge):
  """(u, u*z) of /root/reference/src/pjz/_epsilon.py:47-66,80-85 as float64 tensors of shape
  (2, ll, zz) [z staggering, layer, cell]; differentiable in ``pos`` (the path of the
  layer_pos gradient: the CUDA backward returns d loss / d (u, u*z))."""
  pos = pos.to(torch.float64)
  gs, ge = gs.to(torch.float64).t()[:, None, :], ge.to(torch.float64).t()[:, None, :]   # (2,1,zz)
  inf = torch.full((1,), float("inf"), dtype=torch.float64, device=pos.device)
  lo, hi = torch.cat([-inf, pos])[None, :, None], torch.cat([pos, inf])[None, :, None]
  p0 = torch.minimum(torch.maximum(lo, gs), ge)
  p1 = torch.minimum(torch.maximum(hi, gs), ge)
  u = (p1 - p0) / (ge - gs)
  return u, u * (0.5 * (p0 + p1) - 0.5 * (gs + ge))


def _bind(L):
  L.b200fdtd_render_workspace_bytes.restype = ctypes.c_size_t
  L.b200fdtd_render_backward_workspace_bytes.restype = ctypes.c_size_t
  L.b200fdtd_render.restype = ctypes.c_int
  L.b200fdtd_render.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 4 + [ctypes.c_int] + \
      [ctypes.c_void_p] * 3
  L.b200fdtd_render_backward.restype = ctypes.c_int
  L.b200fdtd_render_backward.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 4 + \
      [ctypes.c_int] + [ctypes.c_void_p] * 5
  return L


class _Render(torch.autograd.Function):
  """b200fdtd_render with its vector-Jacobian product b200fdtd_render_backward: gradients with
  respect to ``layers`` and ``layer_pos`` (what the reference's test differentiates,
  /root/reference/tests/test_layers.py:179-188)."""

  @staticmethod
  def forward(ctx, lay, pos, gs, ge, m, simple):
    from . import fdtdz_jax as shim
    dev = lay.device
    ll, xx, yy = lay.shape[0], lay.shape[1] // (2 * m), lay.shape[2] // (2 * m)
    zz = gs.shape[0]
    L = _bind(shim.lib())
    ws = torch.empty(L.b200fdtd_render_workspace_bytes(ll, xx, yy, zz), dtype=torch.uint8, device=dev)
    out = torch.empty((3, xx, yy, zz), dtype=torch.float32, device=dev)
    pos_arg = pos if pos.numel() else torch.zeros(1, dtype=torch.float32, device=dev)  # never read
    with torch.cuda.device(dev):
      rc = L.b200fdtd_render(ll, xx, yy, zz, m, lay.data_ptr(), pos_arg.data_ptr(), gs.data_ptr(),
                             ge.data_ptr(), int(simple), ws.data_ptr(), out.data_ptr(),
                             torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
      raise RuntimeError(f"b200fdtd_render failed ({rc}): {shim._last_error()}")
    ctx.save_for_backward(lay, pos, gs, ge)
    ctx.m, ctx.simple = m, simple
    return out

  @staticmethod
  def backward(ctx, g):
    from . import fdtdz_jax as shim
    lay, pos, gs, ge = ctx.saved_tensors
    m, simple = ctx.m, ctx.simple
    dev = lay.device
    ll, xx, yy = lay.shape[0], lay.shape[1] // (2 * m), lay.shape[2] // (2 * m)
    zz = gs.shape[0]
    L = _bind(shim.lib())
    g = g.to(torch.float32).contiguous()
    ws = torch.empty(L.b200fdtd_render_backward_workspace_bytes(ll, xx, yy, zz), dtype=torch.uint8,
                     device=dev)
    dlay = torch.empty_like(lay)
    dtab = torch.empty((2, 2, ll, zz), dtype=torch.float64, device=dev)
    pos_arg = pos if pos.numel() else torch.zeros(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
      rc = L.b200fdtd_render_backward(ll, xx, yy, zz, m, lay.data_ptr(), pos_arg.data_ptr(),
                                      gs.data_ptr(), ge.data_ptr(), int(simple), g.data_ptr(),
                                      ws.data_ptr(), dlay.data_ptr(), dtab.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
      raise RuntimeError(f"b200fdtd_render_backward failed ({rc}): {shim._last_error()}")
    dpos = None
    if ctx.needs_input_grad[1] and pos.numel():
      with torch.enable_grad():
        p = pos.detach().to(torch.float64).requires_grad_(True)
        u, uz = _overlap_table(p, gs, ge)
        dpos, = torch.autograd.grad((u * dtab[0]).sum() + (uz * dtab[1]).sum(), p)
      dpos = dpos.to(pos.dtype)
    return (dlay if ctx.needs_input_grad[0] else None), dpos, None, None, None, None


def render(layers, layer_pos, grid_start, grid_end, m, use_simple_averaging=False, device=None):
  """``pjz.render`` (:95-101): ``layers (ll, 2m*xx, 2m*yy)``, ``layer_pos (ll-1,)``,
  ``grid_start`` / ``grid_end (zz, 2)`` -> ``(3, xx, yy, zz)`` float32 CUDA tensor.
  Differentiable with respect to ``layers`` and ``layer_pos`` when they are tensors that require
  grad (``torch.autograd``; the reference uses ``jax.grad``)."""
  if not torch.cuda.is_available():
    raise RuntimeError("render needs a CUDA device (no CPU fallback)")
  dev = torch.device(device) if device is not None else (
      layers.device if isinstance(layers, torch.Tensor) and layers.is_cuda
      else torch.device("cuda", torch.cuda.current_device()))
  lay = _dev_f32(layers, dev)
  m = int(m)
  if lay.ndim != 3 or lay.shape[1] % (2 * m) or lay.shape[2] % (2 * m):
    raise ValueError(f"layers must be (ll, 2m*xx, 2m*yy) with m={m}, got {tuple(lay.shape)}")
  ll = lay.shape[0]
  pos = _dev_f32(np.asarray(layer_pos, np.float32).reshape(-1) if not isinstance(layer_pos, torch.Tensor)
                 else layer_pos.reshape(-1), dev)
  if pos.numel() != ll - 1:
    raise ValueError(f"layer_pos must have {ll - 1} entries, got {pos.numel()}")
  gs, ge = _dev_f32(grid_start, dev), _dev_f32(grid_end, dev)
  if gs.ndim != 2 or gs.shape[1] != 2 or ge.shape != gs.shape:
    raise ValueError("grid_start / grid_end must have shape (zz, 2)")
  return _Render.apply(lay, pos, gs, ge, m, bool(use_simple_averaging))


def epsilon(layers, interface_positions, magnification, zz, use_simple_averaging=False,
            device=None):
  """``pjz.epsilon`` (:104-155): unit cells along z, Ex/Ey centred on integer z, Ez on z+1/2."""
  z = np.arange(int(zz), dtype=np.float32)[:, None]
  return render(layers, interface_positions, z + np.array([[-0.5, 0.0]], np.float32),
                z + np.array([[0.5, 1.0]], np.float32), magnification, use_simple_averaging,
                device=device)
