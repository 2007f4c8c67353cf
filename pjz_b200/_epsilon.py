"""Permittivity on the Yee cell, rendered on the GPU.

Mirrors ``pjz.render`` / ``pjz.epsilon`` (/root/reference/src/pjz/_epsilon.py:95-155): same
arguments, same ``(3, xx, yy, zz)`` result, produced by ``b200fdtd_render`` (csrc/render.cuh)
as a float32 CUDA tensor in the layout ``fdtdz_jax.fdtdz`` takes as ``epsilon`` -- no host
round trip between the renderer and the engine.  Forward only: the reference obtains gradients
with ``jax.grad`` (tests/test_layers.py:191-200), which is outside the time-stepping path.
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


def render(layers, layer_pos, grid_start, grid_end, m, use_simple_averaging=False, device=None):
  """``pjz.render`` (:95-101): ``layers (ll, 2m*xx, 2m*yy)``, ``layer_pos (ll-1,)``,
  ``grid_start`` / ``grid_end (zz, 2)`` -> ``(3, xx, yy, zz)`` float32 CUDA tensor."""
  from . import fdtdz_jax as shim
  if not torch.cuda.is_available():
    raise RuntimeError("render needs a CUDA device (no CPU fallback)")
  dev = torch.device(device) if device is not None else (
      layers.device if isinstance(layers, torch.Tensor) and layers.is_cuda
      else torch.device("cuda", torch.cuda.current_device()))
  lay = _dev_f32(layers, dev)
  m = int(m)
  if lay.ndim != 3 or lay.shape[1] % (2 * m) or lay.shape[2] % (2 * m):
    raise ValueError(f"layers must be (ll, 2m*xx, 2m*yy) with m={m}, got {tuple(lay.shape)}")
  ll, xx, yy = lay.shape[0], lay.shape[1] // (2 * m), lay.shape[2] // (2 * m)
  pos = _dev_f32(np.asarray(layer_pos, np.float32).reshape(-1) if not isinstance(layer_pos, torch.Tensor)
                 else layer_pos.reshape(-1), dev)
  if pos.numel() != ll - 1:
    raise ValueError(f"layer_pos must have {ll - 1} entries, got {pos.numel()}")
  gs, ge = _dev_f32(grid_start, dev), _dev_f32(grid_end, dev)
  if gs.ndim != 2 or gs.shape[1] != 2 or ge.shape != gs.shape:
    raise ValueError("grid_start / grid_end must have shape (zz, 2)")
  zz = gs.shape[0]
  L = shim.lib()
  L.b200fdtd_render_workspace_bytes.restype = ctypes.c_size_t
  L.b200fdtd_render.restype = ctypes.c_int
  L.b200fdtd_render.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 4 + [ctypes.c_int] + \
      [ctypes.c_void_p] * 3
  ws = torch.empty(L.b200fdtd_render_workspace_bytes(ll, xx, yy, zz), dtype=torch.uint8, device=dev)
  out = torch.empty((3, xx, yy, zz), dtype=torch.float32, device=dev)
  if pos.numel() == 0:
    pos = torch.zeros(1, dtype=torch.float32, device=dev)      # never read (ll == 1)
  with torch.cuda.device(dev):
    rc = L.b200fdtd_render(ll, xx, yy, zz, m, lay.data_ptr(), pos.data_ptr(), gs.data_ptr(),
                           ge.data_ptr(), int(bool(use_simple_averaging)), ws.data_ptr(),
                           out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
  if rc != 0:
    raise RuntimeError(f"b200fdtd_render failed ({rc}): {shim._last_error()}")
  return out


def epsilon(layers, interface_positions, magnification, zz, use_simple_averaging=False,
            device=None):
  """``pjz.epsilon`` (:104-155): unit cells along z, Ex/Ey centred on integer z, Ez on z+1/2."""
  z = np.arange(int(zz), dtype=np.float32)[:, None]
  return render(layers, interface_positions, z + np.array([[-0.5, 0.0]], np.float32),
                z + np.array([[0.5, 1.0]], np.float32), magnification, use_simple_averaging,
                device=device)
