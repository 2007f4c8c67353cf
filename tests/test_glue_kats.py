"""Known-answer tests of the host glue, taken from the reference's own tests.

These are the only numbers the reference pins next to the engine boundary:
  absorption mask   /root/reference/tests/test_boundaries.py:8-24
  PML sigma         /root/reference/tests/test_boundaries.py:27-41
  sampling interval /root/reference/tests/test_frequencies.py:8-18
  ramped sine       /root/reference/tests/test_waveforms.py:8-14
"""

import numpy as np
import pytest

from pjz_b200 import _field as glue


def test_absorption_mask():
  m = glue._absorption_mask(xx=10, yy=10, width=3, smoothness=4)
  assert m.shape == (3, 10, 10)
  np.testing.assert_array_equal(m[0, :, 5], [36, 16, 4, 0, 0, 0, 0, 4, 16, 36])
  np.testing.assert_array_equal(m[0, 5, :], [25, 9, 1, 0, 0, 0, 1, 9, 25, 49])
  np.testing.assert_array_equal(m[1, :, 5], [25, 9, 1, 0, 0, 0, 1, 9, 25, 49])
  np.testing.assert_array_equal(m[1, 5, :], [36, 16, 4, 0, 0, 0, 0, 4, 16, 36])
  np.testing.assert_array_equal(m[2, :, 5], [25, 9, 1, 0, 0, 0, 1, 9, 25, 49])
  np.testing.assert_array_equal(m[2, 5, :], [25, 9, 1, 0, 0, 0, 1, 9, 25, 49])


def test_pml_sigma():
  np.testing.assert_array_almost_equal(
      glue._pml_sigma(pml_widths=(4, 4), zz=10, ln_R=16.0, m=4.0),
      [[8.0000000e+01, 4.6894531e+01],
       [2.5312500e+01, 1.2207031e+01],
       [5.0000000e+00, 1.5820312e+00],
       [3.1250000e-01, 1.9531250e-02],
       [0.0000000e+00, 0.0000000e+00],
       [0.0000000e+00, 0.0000000e+00],
       [1.9531250e-02, 3.1250000e-01],
       [1.5820312e+00, 5.0000000e+00],
       [1.2207031e+01, 2.5312500e+01],
       [4.6894531e+01, 8.0000000e+01]])


def test_pml_sigma_zero_width():
  s = glue._pml_sigma(pml_widths=(0, 4), zz=10, ln_R=16.0, m=4.0)
  assert np.all(s[:6] == 0) and s[-1, 1] == 80.0


def test_sampling_interval():
  wmin, wmax = 2 * np.pi / 40, 2 * np.pi / 36
  n, dt = 10, 0.5
  interval = glue._sampling_interval(wmin, wmax, n, dt)
  assert interval == 322  # SURVEY.md 4 [PROBED]
  ws = np.linspace(wmin, wmax, n)
  theta = ws * dt * interval * np.arange(2 * n)[:, None]
  A = np.hstack([np.sin(theta), np.cos(theta)])
  A /= np.linalg.norm(A, ord=2, axis=0)
  assert np.all(A.T @ A - np.eye(2 * n) < 1e-1)


def test_sampling_interval_single_frequency():
  # quarter period of omega = 2 pi / 37 at dt = 0.5 -> 74 / 4 = 18.5 -> 18
  w = 2 * np.pi / 37
  assert glue._sampling_interval(w, w, 1, 0.5) == int(round(37 / 0.5 / 4))


def test_ramped_sin():
  omega, dt, tt = 2 * np.pi / 40, 0.5, 100000
  out = glue._ramped_sin(np.array([omega]), width=3, delay=4, dt=dt, tt=tt)
  np.testing.assert_array_almost_equal(out.imag[-10:], np.sin(omega * dt * np.arange(tt))[-10:])
  np.testing.assert_array_almost_equal(out.real[-10:], np.cos(omega * dt * np.arange(tt))[-10:])
  assert abs(out[0]) < 1e-3  # starts from rest


def test_output_phases_and_projection_recover_phasor():
  """Snapshots of Re(a e^{+i w t}) (the sign convention of the cos / -sin rows,
  /root/reference/src/pjz/_field.py:142-150, 276-279) project back onto a."""
  import torch
  omega = np.array([2 * np.pi / 40, 2 * np.pi / 37])
  dt, tt = 0.5, 4000
  interval = glue._sampling_interval(omega.min(), omega.max(), 2, dt)
  steps = (tt - 2 * interval * 2 - 1, tt, interval)
  n = np.arange(*steps)
  assert len(n) == 5
  a = np.array([0.7 - 0.2j, -0.3 + 1.1j])
  snaps = sum(np.real(a[i] * np.exp(1j * omega[i] * dt * n)) for i in range(2))
  fields = np.broadcast_to(snaps[:, None, None, None, None], (5, 3, 1, 1, 2)).copy()
  out = glue.project_snapshots(torch.from_numpy(fields.astype(np.float32)), omega, steps, dt)
  np.testing.assert_allclose(out[:, 0, 0, 0, 0].numpy(), a, rtol=2e-4, atol=2e-4)


def test_zz_and_padding_rules():
  # /root/reference/src/pjz/_field.py:56-66
  assert glue._zz((16, 16), True) == 96 and glue._zz((8, 8), False) == 48
  assert glue._pad_zz(64, (16, 16), True) == (16, 16)
  assert glue._pad_zz(63, (16, 16), True) == (16, 17)
  assert glue._zz((16, 16), False, domain_zz=128) == 128
  with pytest.raises(ValueError):
    glue._pad_zz(100, (16, 16), True)


def test_simparams_defaults_match_reference():
  # /root/reference/src/pjz/_field.py:40-53
  p = glue.SimParams(omega_range=(0.1, 0.2), tt=100)
  assert (p.dt, p.source_ramp, p.source_delay, p.absorption_padding, p.absorption_coeff,
          p.pml_widths, p.pml_alpha_coeff, p.pml_sigma_lnr, p.pml_sigma_m, p.use_z_as_batch,
          p.use_reduced_precision, p.launch_params) == (
              0.5, 4.0, 4.0, 50, 1e-4, (16, 16), 0.0, 0.5, 1.3, False, True, None)
  assert p._fields[:14] == (
      "omega_range", "tt", "dt", "source_ramp", "source_delay", "absorption_padding",
      "absorption_coeff", "pml_widths", "pml_alpha_coeff", "pml_sigma_lnr", "pml_sigma_m",
      "use_z_as_batch", "use_reduced_precision", "launch_params")
