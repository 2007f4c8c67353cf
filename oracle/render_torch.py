"""Differentiable float64 restatement of pjz's permittivity renderer in torch (CPU).
TEST INFRASTRUCTURE ONLY -- the checker of the CUDA renderer's backward pass.

Follows /root/reference/src/pjz/_epsilon.py:10-101 with the same operations as
oracle/render_numpy.py (which tests/test_render.py pins against every golden value of
/root/reference/tests/test_layers.py); tests check that the two agree in the forward direction
and that this one's autograd gradient matches central differences of the NumPy oracle, before
using it to check ``b200fdtd_render_backward`` at larger sizes.  The reference differentiates
the renderer with ``jax.grad`` (tests/test_layers.py:179-188).
"""

import torch
import torch.nn.functional as F


def _render_single(layers, layer_pos, grid_start, grid_end, m, axis, use_simple_averaging):
  if axis != "x":                                            # in-plane offsets (:13-17)
    layers = torch.cat([layers[:, :1, :].expand(-1, m, -1), layers[:, :-m, :]], 1)
  if axis != "y":
    layers = torch.cat([layers[:, :, :1].expand(-1, -1, m), layers[:, :, :-m]], 2)
  col = 1 if axis == "z" else 0                              # offsets along z (:19-25)
  gs, ge = grid_start[:, col], grid_end[:, col]
  lc = layers.reshape(layers.shape[0], layers.shape[1] // (2 * m), 2 * m,
                      layers.shape[2] // (2 * m), 2 * m)     # "layer-chunked" form (:27-30)
  w = (torch.arange(2 * m, dtype=layers.dtype) - (m - 0.5)) / (2 * m)**2   # (:32-33)
  grads = [torch.mean(12 * (2 * m) * x, (2, 4)) for x in (lc * w[:, None, None], lc * w)]
  avg = torch.mean(lc, (2, 4))
  aoi = torch.mean(1 / lc, (2, 4))
  inf = torch.full((1,), float("inf"), dtype=layers.dtype)
  pos = layer_pos.reshape(-1)
  lo, hi = torch.cat([-inf, pos])[:, None], torch.cat([pos, inf])[:, None]
  p0 = torch.minimum(torch.maximum(lo, gs), ge)              # (:55-58)
  p1 = torch.minimum(torch.maximum(hi, gs), ge)
  u = (p1 - p0) / (ge - gs)
  cross = lambda x, y: torch.einsum("lxy,lz->xyz", x, y)
  if use_simple_averaging:
    return cross(avg, u)
  z = (p0 + p1) / 2 - (gs + ge) / 2
  aoi = cross(aoi, u)
  ioa = 1 / cross(avg, u)
  grads = [cross(g, u) for g in grads]
  grads.append(cross(avg, u * z) / ((ge - gs)**2 / 12))
  ssq = sum(g**2 for g in grads)
  pii = grads["xyz".index(axis)]**2 / torch.where(ssq == 0, torch.ones_like(ssq), ssq)
  return 1 / (pii * aoi + (1 - pii) * ioa)


def render(layers, layer_pos, grid_start, grid_end, m, use_simple_averaging=False):
  """float64 tensors in, ``(3, xx, yy, zz)`` float64 out; autograd-differentiable."""
  args = [torch.as_tensor(a, dtype=torch.float64) if not isinstance(a, torch.Tensor) else a.double()
          for a in (layers, layer_pos, grid_start, grid_end)]
  return torch.stack([_render_single(*args, m, axis, use_simple_averaging) for axis in "xyz"])
