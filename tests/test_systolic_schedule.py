"""The systolic kernel's dependency rule, checked on the CPU by randomised interleaving
(tests/systolic_emulator.py).  With the rule the kernel uses (`k+3`) every interleaving must
reproduce the oracle exactly; a weaker rule must be caught."""

import numpy as np
import pytest

from oracle import fdtd_numpy
from tests.problems import random_problem
from tests.systolic_emulator import Emulator


@pytest.mark.parametrize("ntiles,stages,axis,seed", [
    (3, 4, 0, 0), (2, 3, 1, 1), (1, 5, 2, 2), (4, 2, 0, 3), (5, 7, 1, 4), (1, 1, 2, 5)])
def test_schedule_is_exact_under_random_interleaving(ntiles, stages, axis, seed):
  kw = random_problem(domain=(9, 11, 8), axis=axis, tt=11, seed=seed, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  out = Emulator(kw, ntiles, stages, max_lead=4, need_rule=3, seed=seed).run()
  np.testing.assert_array_equal(out, ref)


def test_tiny_domains_wrap_correctly():
  for dom in [(1, 4, 4), (2, 3, 4), (3, 1, 4), (4, 2, 1)]:
    kw = random_problem(domain=dom, sub=dom, offset=(0, 0, 0), axis=0, pml=(0, 0), tt=7, seed=1,
                        output_steps=(0, 7, 1), absorb_pad=0)
    ref = fdtd_numpy.fdtdz(**kw)
    out = Emulator(kw, min(2, dom[1]), 3, max_lead=4, seed=2).run()
    np.testing.assert_array_equal(out, ref)


def test_weaker_rule_is_detected():
  kw = random_problem(domain=(9, 11, 8), axis=0, tt=11, seed=0, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  out = Emulator(kw, 3, 4, max_lead=4, need_rule=2, seed=0).run()
  assert not np.array_equal(out, ref)


@pytest.mark.parametrize("ntiles,stages,axis,seed", [
    (2, 4, 0, 10), (3, 3, 1, 11), (1, 5, 2, 12), (4, 2, 0, 13), (2, 7, 1, 14), (3, 6, 2, 15)])
def test_l2_discard_rule_never_drops_live_data(ntiles, stages, axis, seed):
  """kernels_lean.cuh drops the field lines a tile has just consumed (discard.global.L2) on its
  exclusive columns, except the sweep's first loads.  Modelled as NaN poisoning under random
  interleaving: the result must still be exact, i.e. nothing poisoned is ever read again before
  it is rewritten."""
  kw = random_problem(domain=(9, 23, 8), axis=axis, tt=13, seed=seed, output_steps=(3, 13, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  out = Emulator(kw, ntiles, stages, max_lead=4, need_rule=3, seed=seed, discard=True).run()
  assert np.isfinite(out).all()
  np.testing.assert_array_equal(out, ref)


def test_discarding_the_first_loads_of_a_sweep_is_detected():
  """The planes a sweep loads first are loaded AGAIN when the sweep wraps around the periodic x
  boundary: dropping them after the first load (the exemption removed) must corrupt the result."""
  kw = random_problem(domain=(9, 23, 8), axis=0, tt=13, seed=10, output_steps=(3, 13, 2))
  out = Emulator(kw, 2, 4, max_lead=4, need_rule=3, seed=10, discard=True, discard_first=True).run()
  assert not np.isfinite(out).all()
