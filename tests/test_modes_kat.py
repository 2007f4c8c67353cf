"""Mode-solver known answers from /root/reference/tests/test_modes.py."""

import numpy as np
import pytest

from pjz_b200 import _mode as M
from pjz_b200 import mode


def _eps(prop_axis, uu=30, vv=20):
  eps = np.ones((3, uu, vv))
  eps[:, 9:21, 8:12] = 12.25
  return np.expand_dims(eps, axis="xyz".find(prop_axis) + 1)


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_mode_output_is_float32_and_correct_shape(prop_axis):
  # /root/reference/tests/test_modes.py:7-31
  uu, vv = 30, 20
  omega = np.linspace(2 * np.pi / 37, 2 * np.pi / 31, 2)
  beta, field, err, iters = mode(_eps(prop_axis), omega, num_modes=3)
  assert beta.dtype == np.float32 and field.dtype == np.float32 and err.dtype == np.float32
  assert beta.shape == (2, 3) and err.shape == (2, 3)
  assert field.shape == {"x": (2, 2, 1, uu, vv, 3), "y": (2, 2, uu, 1, vv, 3),
                         "z": (2, 2, uu, vv, 1, 3)}[prop_axis]


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_correct_betas(prop_axis):
  # /root/reference/tests/test_modes.py:34-47
  expected = (0.36388508, 0.18891069, 0.15406249, 0.13549446)
  beta, _, _, _ = mode(_eps(prop_axis), np.array([2 * np.pi / 37]), num_modes=4)
  assert beta[0, :] == pytest.approx(expected, rel=1e-3)
  assert beta[0, :] == pytest.approx(expected, rel=1e-6)  # reproduces to 8 digits


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_full_fields_self_consistent_and_unit_poynting(prop_axis):
  # /root/reference/tests/test_modes.py:76-115 (same un-permutation of the excitation)
  ww, mm = 2, 2
  omega = np.linspace(2 * np.pi / 37, 2 * np.pi / 36, ww)
  epsilon = _eps(prop_axis)
  beta, field, _, _ = mode(epsilon, omega, num_modes=mm)
  if prop_axis == "x":
    f = np.flip(field[:, :, 0, :, :, :], axis=1)
    epsilon = epsilon[(1, 2, 0), ...]
  elif prop_axis == "y":
    f = np.flip(np.swapaxes(field[:, :, :, 0, :, :], 2, 3), axis=2)
    epsilon = np.flip(np.swapaxes(epsilon[(2, 0, 1), ...], 1, 3), axis=1)
  else:
    f = np.array([-1, 1])[None, :, None, None, None] * np.flip(field[:, :, :, :, 0, :], axis=1)
  eps2 = np.squeeze(epsilon)
  for w in range(ww):
    for k in range(mm):
      x = f[w, ..., k].astype(np.float64)
      h, e, h2 = M._full_fields(float(beta[w, k]), omega[w], eps2, x)
      assert np.linalg.norm(h - h2) / np.linalg.norm(h) < 1e-4   # reference bound: 1e-2
      p = np.sum(e[0] * h[1] - e[1] * h[0])
      assert abs(p - 1) < 1e-5


def test_rejects_non_singleton():
  with pytest.raises(ValueError):
    mode(np.ones((3, 4, 4, 4)), np.array([0.2]), 1)
