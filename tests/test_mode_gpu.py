"""GPU mode solver (pjz_b200/_mode_gpu.py) against the reference's own
known answers (/root/reference/tests/test_modes.py) and the host harness (pjz_b200/_mode.py)."""

import numpy as np
import pytest
import torch

from pjz_b200 import _mode as M

pytestmark = pytest.mark.gpu


def _eps(prop_axis, uu=30, vv=20):
  eps = np.ones((3, uu, vv))
  eps[:, 9:21, 8:12] = 12.25
  return np.expand_dims(eps, axis="xyz".find(prop_axis) + 1)


def test_operator_kernel_matches_the_reference_operator(built):
  """b200fdtd_mode_operator vs the NumPy restatement of _mode.py:22-51, float32 tolerance."""
  from pjz_b200 import _mode_gpu as G
  rng = np.random.default_rng(0)
  ww, uu, vv, mm = 3, 17, 11, 4
  eps = rng.uniform(1, 12, (3, uu, vv))
  omega = rng.uniform(0.1, 0.3, ww)
  shift = rng.uniform(-1, 1, ww)
  x = rng.standard_normal((ww, 2, uu, vv, mm))
  want = np.stack([np.stack([M._apply_operator(eps, omega[w], x[w, ..., m]) - shift[w] * x[w, ..., m]
                             for m in range(mm)], axis=-1) for w in range(ww)])
  t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
  got = G._apply(t(eps), t(omega), t(shift), t(x)).cpu().numpy()
  assert np.linalg.norm(got - want) <= 2e-6 * np.linalg.norm(want)


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_correct_betas(prop_axis, built):
  # /root/reference/tests/test_modes.py:34-47
  from pjz_b200._mode_gpu import mode_gpu
  expected = (0.36388508, 0.18891069, 0.15406249, 0.13549446)
  beta, field, err, iters = mode_gpu(_eps(prop_axis), np.array([2 * np.pi / 37]), num_modes=4)
  assert beta.is_cuda and beta.dtype == torch.float32 and field.dtype == torch.float32
  assert beta[0, :].cpu().numpy() == pytest.approx(expected, rel=1e-3)
  uu, vv = 30, 20
  assert tuple(field.shape) == {"x": (1, 2, 1, uu, vv, 4), "y": (1, 2, uu, 1, vv, 4),
                                "z": (1, 2, uu, vv, 1, 4)}[prop_axis]
  assert float(err.max()) <= 1e-4 and iters >= 1


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_warm_start_converges_in_one_iteration(prop_axis, built):
  # /root/reference/tests/test_modes.py:50-73
  from pjz_b200._mode_gpu import mode_gpu
  omega = np.array([2 * np.pi / 37])
  beta, field, _, _ = mode_gpu(_eps(prop_axis), omega, num_modes=1)
  beta2, _, _, iters = mode_gpu(_eps(prop_axis), omega, num_modes=1, init=field)
  np.testing.assert_array_almost_equal(beta.cpu().numpy(), beta2.cpu().numpy(), decimal=3)
  assert iters == 1


@pytest.mark.parametrize("prop_axis", ["x", "y", "z"])
def test_unit_poynting_self_consistency_and_agreement_with_the_host_harness(prop_axis, built):
  # /root/reference/tests/test_modes.py:76-115, batched over 3 frequencies
  from pjz_b200 import mode
  from pjz_b200._mode_gpu import mode_gpu
  ww, mm = 3, 2
  omega = np.linspace(2 * np.pi / 37, 2 * np.pi / 36, ww)
  epsilon = _eps(prop_axis)
  beta, field, _, _ = mode_gpu(epsilon, omega, num_modes=mm)
  beta, field = beta.cpu().numpy(), field.cpu().numpy()
  host_beta, _, _, _ = mode(epsilon, omega, num_modes=mm)
  assert beta == pytest.approx(host_beta, rel=1e-3)
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
      assert np.linalg.norm(h - h2) / np.linalg.norm(h) < 1e-2       # the reference's bound
      assert abs(np.sum(e[0] * h[1] - e[1] * h[0]) - 1) < 1e-3


def test_rejects_non_singleton(built):
  from pjz_b200._mode_gpu import mode_gpu
  with pytest.raises(ValueError):
    mode_gpu(np.ones((3, 4, 4, 4)), np.array([0.2]), 1)


# ---- /root/reference/tests/test_integration.py: renderer + mode solver under sub-cell shifts ---------

def _rect(pos, center, widths):
  """pjz.rect as documented in /root/reference/src/pjz/_shape.py:16-17."""
  out = 1.0
  for p, c, w in zip(pos, center, widths):
    out = out * np.clip((w + 1) / 2 - np.abs(p - c), 0, 1)
  return out


def test_beta_spread_under_subcell_shifts_xy(built):
  # test_integration.py:8-28: a 25x10 Si block shifted by fifths of a cell in x and y
  from pjz_b200._epsilon import render
  from pjz_b200._mode_gpu import mode_gpu
  xx, yy = 30, 20
  pos = (np.arange(2 * xx)[:, None], np.arange(2 * yy))
  betas = []
  for x in np.arange(0, 1, 0.2):
    for y in np.arange(0, 1, 0.2):
      eps = 1 + 12.25 * _rect(pos, (xx + x, yy + y), (25, 10))
      epsilon = render(eps[None, :, :].astype(np.float32), np.array([]), np.zeros((1, 2)),
                       np.ones((1, 2)), 1)
      beta, _, _, _ = mode_gpu(epsilon, np.array([2 * np.pi / 37]), 1)
      betas.append(float(beta[0, 0]))
  assert (np.max(betas) - np.min(betas)) / 2 / np.mean(betas) <= 1e-2


def test_beta_spread_under_subcell_shifts_xz(built):
  # test_integration.py:31-49: a slab whose z interfaces move by fifths of a cell
  from pjz_b200._epsilon import render
  from pjz_b200._mode_gpu import mode_gpu
  xx, yy, zz = 30, 1, 20
  pos = (np.arange(2 * xx)[:, None], np.arange(2 * yy))
  betas = []
  for x in np.arange(0, 1, 0.2):
    for z in np.arange(0, 1, 0.2):
      eps = np.ones((3, 2 * xx, 2 * yy))
      eps[1, ...] = 1 + 12.25 * _rect(pos, (xx + x, yy), (25, np.inf))
      epsilon = render(eps.astype(np.float32), np.array([7.5, 12.5]) + z,
                       np.arange(zz)[:, None] + np.array([[-0.5, 0]]),
                       np.arange(zz)[:, None] + np.array([[0.5, 1.0]]), 1)
      assert tuple(epsilon.shape) == (3, xx, 1, zz)
      beta, _, _, _ = mode_gpu(epsilon, np.array([2 * np.pi / 37]), 1)
      betas.append(float(beta[0, 0]))
  assert (np.max(betas) - np.min(betas)) / 2 / np.mean(betas) <= 1e-2


def test_buried_waveguide_cross_section_matches_the_host_harness(built):
  """A cfg3-style port (Si core in a 2.25 cladding, 96x48 cells, 2 frequencies, 2 modes): betas
  equal to the ARPACK harness, residuals below tol."""
  from pjz_b200 import mode
  from pjz_b200._mode_gpu import mode_gpu
  uu, vv = 96, 48
  eps = np.full((3, 1, uu, vv), 2.25, np.float32)
  eps[:, :, uu // 2 - 6:uu // 2 + 6, vv // 2 - 4:vv // 2 + 4] = 12.25
  omega = np.array([2 * np.pi / 40, 2 * np.pi / 36])
  beta, exc, err, iters = mode_gpu(eps, omega, 2)
  host_beta, _, _, _ = mode(eps, omega, 2)
  assert torch.isfinite(exc).all() and float(err.max()) <= 1e-4 and iters < 100
  assert beta.cpu().numpy() == pytest.approx(host_beta, rel=1e-4)
