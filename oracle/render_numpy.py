"""NumPy restatement of pjz's permittivity renderer.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/pjz/_epsilon.py:10-101 (``_render_single`` / ``_render``) line by
line in float64.  PINNED: tests/test_render.py checks it against every golden value of
/root/reference/tests/test_layers.py; the CUDA renderer (csrc/render.cuh) is then compared with
it on random inputs.
"""

import numpy as np


def _render_single(layers, layer_pos, grid_start, grid_end, m, axis, use_simple_averaging):
  layers = np.asarray(layers, np.float64)
  if axis != "x":                                            # in-plane offsets (:13-17)
    layers = np.pad(layers[:, :-m, :], ((0, 0), (m, 0), (0, 0)), "edge")
  if axis != "y":
    layers = np.pad(layers[:, :, :-m], ((0, 0), (0, 0), (m, 0)), "edge")
  col = 1 if axis == "z" else 0                              # offsets along z (:19-25)
  gs, ge = np.asarray(grid_start, np.float64)[:, col], np.asarray(grid_end, np.float64)[:, col]
  lc = layers.reshape(layers.shape[0], layers.shape[1] // (2 * m), 2 * m,
                      layers.shape[2] // (2 * m), 2 * m)     # "layer-chunked" form (:27-30)
  w = (np.arange(2 * m) - (m - 0.5)) / (2 * m)**2            # (:32-33)
  grads = [np.mean(12 * (2 * m) * x, (2, 4)) for x in (lc * w[:, None, None], lc * w)]
  avg = np.mean(lc, (2, 4))
  aoi = np.mean(1 / lc, (2, 4))
  pos = np.asarray(layer_pos, np.float64).reshape(-1)
  p0, p1 = [np.clip(x[:, None], gs, ge) for x in
            (np.concatenate([[-np.inf], pos]), np.concatenate([pos, [np.inf]]))]   # (:55-58)
  u = (p1 - p0) / (ge - gs)
  cross = lambda x, y: np.einsum("lxy,lz->xyz", x, y)
  if use_simple_averaging:
    return cross(avg, u)
  z = (p0 + p1) / 2 - (gs + ge) / 2
  aoi = cross(aoi, u)
  ioa = 1 / cross(avg, u)
  grads = [cross(g, u) for g in grads]
  grads.append(cross(avg, u * z) / ((ge - gs)**2 / 12))
  ssq = sum(g**2 for g in grads)
  pii = grads["xyz".index(axis)]**2 / np.where(ssq == 0, 1, ssq)
  return 1 / (pii * aoi + (1 - pii) * ioa)


def render(layers, layer_pos, grid_start, grid_end, m, use_simple_averaging=False):
  return np.stack([_render_single(layers, layer_pos, grid_start, grid_end, m, axis,
                                  use_simple_averaging) for axis in "xyz"])
