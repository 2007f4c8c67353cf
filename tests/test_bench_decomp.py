"""bench.py's decomposed-domain mode builds every rank's slab of the cfg5 metalens directly (the
4096x4096x128 stack is never materialised).  That shortcut must give exactly what the generic
decomposition (`pjz_b200._decomp.local_problem_y`) cuts out of the global problem."""

import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
  spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


@pytest.mark.parametrize("world,ghost", [(1, 1), (2, 1), (2, 8), (4, 1), (4, 16)])
def test_per_rank_metalens_slab_equals_the_cut_of_the_global_problem(world, ghost):
  from pjz_b200 import _field as glue
  from pjz_b200 import workloads as W
  from pjz_b200._decomp import local_problem_y
  bench = _bench()
  total, tt = (80, 96 * world if world < 4 else 256, 128), 12
  X, Y, Z = total
  eps, _, _, _ = W.metalens(total=total)
  loc0, _, _ = bench._metalens_slab(total, 0, world, ghost, tt, "cpu")
  kw = dict(loc0)
  kw["epsilon"] = eps
  kw["absorption_mask"] = glue._absorption_mask(X, Y, 32, 1e-4)
  kw["source_field"] = np.zeros((2, 2, X, Y, 1), np.float32)
  kw["source_field"][1, 0] = 0.01
  kw["offset"] = (32, 32, 16)
  for rank in range(world):
    want, nloc_w, crop_w = local_problem_y(kw, rank, world, ghost)
    got, nloc, crop = bench._metalens_slab(total, rank, world, ghost, tt, "cpu")
    assert nloc == nloc_w and crop == crop_w
    assert tuple(got["offset"]) == tuple(want["offset"])
    np.testing.assert_array_equal(got["epsilon"].numpy(), want["epsilon"])
    np.testing.assert_array_equal(got["absorption_mask"], want["absorption_mask"])
    np.testing.assert_array_equal(got["source_field"].numpy(), want["source_field"])
    assert got["source_position"] == want["source_position"]
    for k in ("pml_kappa", "pml_sigma", "pml_alpha", "source_waveform"):
      np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(want[k]))


def test_metalens_columns_equals_the_numpy_builder():
  from pjz_b200 import workloads as W
  total = (96, 128, 128)
  eps, _, _, _ = W.metalens(total=total)
  cols = np.arange(total[1] - 64)
  np.testing.assert_array_equal(W.metalens_columns(total, cols, "cpu").numpy(), eps)
  sub = np.array([0, 0, 5, 17, 63, 63])
  np.testing.assert_array_equal(W.metalens_columns(total, sub, "cpu").numpy(), eps[:, :, sub])
