"""CPU tests of the oracle itself: the C fp32 restatement against the float64 NumPy spec, both
against the committed golden vectors, and the size-independent properties of the update
(linearity in the source, z-as-batch decoupling, time-translation of the output schedule)."""

import os

import numpy as np
import pytest

from oracle import fdtd_c, fdtd_numpy
from tests.golden.make_golden import CASES
from tests.problems import random_problem, rel_l2

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "engine_golden.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_spec_reproduces_golden(name):
  out = fdtd_numpy.fdtdz(**random_problem(**CASES[name]))
  np.testing.assert_array_equal(out, GOLDEN[name])


@pytest.mark.parametrize("name", sorted(CASES))
def test_c_oracle_matches_golden(name):
  out = fdtd_c.fdtdz(**random_problem(**CASES[name]))
  assert out.shape == GOLDEN[name].shape
  assert rel_l2(out, GOLDEN[name]) < 1e-6


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("pml", [(0, 0), (4, 6)])
def test_c_vs_numpy_stress(axis, pml):
  # SURVEY.md 8(d) "stress": 48x40x32, eps~U[1,12.25], random source, all orientations.
  kw = random_problem(domain=(48, 40, 32), sub=(40, 30, 20), axis=axis, pml=pml, tt=60,
                      seed=2 + axis, output_steps=(30, 60, 7))
  a = fdtd_numpy.fdtdz(**kw)
  b = fdtd_c.fdtdz(**kw)
  assert rel_l2(b, a) < 1e-6


def test_c_threads_do_not_change_bits():
  kw = random_problem(domain=(16, 12, 16), tt=20, seed=9)
  a = fdtd_c.fdtdz(**kw, nthreads=1)
  b = fdtd_c.fdtdz(**kw, nthreads=4)
  np.testing.assert_array_equal(a, b)


def test_reduced_precision_mode():
  kw = random_problem(domain=(12, 10, 16), tt=40, seed=5, reduced=True, output_steps=(20, 40, 5))
  a = fdtd_numpy.fdtdz(**kw)
  b = fdtd_c.fdtdz(**kw)
  full = fdtd_numpy.fdtdz(**{**kw, "use_reduced_precision": False})
  assert rel_l2(b, a) < 2e-3          # two fp16-storage implementations, different op order
  assert 1e-5 < rel_l2(a, full) < 5e-3  # fp16 storage is visible but bounded


def test_linearity_in_source():
  kw = random_problem(domain=(14, 12, 12), tt=30, seed=21, output_steps=(10, 30, 5))
  kw2 = dict(kw)
  kw2["source_field"] = kw["source_field"] * 2.0
  a, b = fdtd_c.fdtdz(**kw), fdtd_c.fdtdz(**kw2)
  assert rel_l2(b, 2 * a) < 1e-6


def test_zero_source_gives_zero():
  kw = random_problem(domain=(8, 8, 8), tt=10, seed=3)
  kw["source_field"] = np.zeros_like(kw["source_field"])
  assert not fdtd_c.fdtdz(**kw).any()


def test_z_as_batch_decouples_planes():
  """kappa = inf (use_z_as_batch, /root/reference/src/pjz/_field.py:243-246): every z-plane is an
  independent 2-D simulation -> running plane z alone reproduces slice z of the 3-D run."""
  kw = random_problem(domain=(12, 10, 6), sub=(12, 10, 6), offset=(0, 0, 0), axis=0, pml=(0, 0),
                      tt=25, seed=8, z_as_batch=True, output_steps=(5, 25, 5))
  full = fdtd_c.fdtdz(**kw)
  for z in (0, 3, 5):
    kw1 = dict(kw)
    kw1["epsilon"] = kw["epsilon"][..., z:z + 1]
    kw1["source_field"] = kw["source_field"][..., z:z + 1]
    for k in ("pml_kappa", "pml_sigma", "pml_alpha"):
      kw1[k] = kw[k][z:z + 1]
    one = fdtd_c.fdtdz(**kw1)
    np.testing.assert_array_equal(one[..., 0], full[..., z])


def test_output_schedule_is_a_pure_selection():
  kw = random_problem(domain=(10, 8, 8), tt=20, seed=4, output_steps=(0, 20, 1))
  every = fdtd_c.fdtdz(**kw)
  kw["output_steps"] = (3, 18, 4)
  some = fdtd_c.fdtdz(**kw)
  np.testing.assert_array_equal(some, every[3:18:4])


def test_epsilon_edge_replication():
  """epsilon is the sub-volume at `offset`; outside it is edge-replicated."""
  kw = random_problem(domain=(10, 9, 8), sub=(4, 3, 2), offset=(3, 2, 4), tt=15, seed=6,
                      output_steps=(14, 15, 1))
  big = dict(kw)
  big["epsilon"] = fdtd_numpy.extend_epsilon(kw["epsilon"], (10, 9, 8), (3, 2, 4))
  big["offset"] = (0, 0, 0)
  a = fdtd_c.fdtdz(**kw)
  b = fdtd_c.fdtdz(**big)
  np.testing.assert_array_equal(a, b[:, :, 3:7, 2:5, 4:6])


def test_bad_shapes_raise():
  kw = random_problem()
  kw["source_field"] = kw["source_field"][:, :, :-1]
  with pytest.raises(ValueError):
    fdtd_c.fdtdz(**kw)
  with pytest.raises(ValueError):
    fdtd_numpy.fdtdz(**kw)


def test_projection_extension_c_vs_spec():
  """``output_projection`` (engine extension): C oracle's fmaf running sum vs the float64 spec's
  einsum over the same snapshots; and it is exactly W @ snapshots in exact arithmetic."""
  from oracle import fdtd_c, fdtd_numpy
  from tests.problems import random_problem
  kw = random_problem(domain=(10, 9, 12), axis=2, pml=(3, 3), tt=20, seed=4, output_steps=(4, 20, 5))
  W = np.random.default_rng(1).standard_normal((3, 4)).astype(np.float32)
  snaps = fdtd_c.fdtdz(**kw)
  got = fdtd_c.fdtdz(**kw, output_projection=W)
  spec = fdtd_numpy.fdtdz(**kw, output_projection=W)
  assert got.shape == spec.shape == (3,) + snaps.shape[1:]
  ref = np.einsum("rs,s...->r...", W.astype(np.float64), snaps.astype(np.float64))
  np.testing.assert_allclose(got, ref, rtol=0, atol=4e-7 * np.abs(ref).max())
  assert np.linalg.norm(got - spec) <= 1e-5 * np.linalg.norm(spec)
  with pytest.raises(ValueError):
    fdtd_c.fdtdz(**kw, output_projection=W[:, :3])
