"""The y-slab sessions' protocol (courier CTAs, mirror slots, the courier's write-after-read rule;
pjz_b200/csrc/kernels_lean.cuh, SLAB = true), checked on the CPU by randomised interleaving of
tiles AND couriers (tests/slab_emulator.py): every interleaving must reproduce the single-domain
oracle exactly; without the write-after-read rule some interleaving must not."""

import numpy as np
import pytest

from oracle import fdtd_numpy
from tests.problems import random_problem
from tests.slab_emulator import SlabEmulator


@pytest.mark.parametrize("world,tiles,stages,axis,seed", [
    (2, 2, 4, 0, 0), (2, 1, 3, 1, 1), (1, 3, 5, 2, 2), (3, 1, 2, 0, 3), (2, 3, 7, 1, 4),
    (1, 1, 4, 2, 5), (4, 1, 3, 2, 6), (2, 2, 1, 0, 7)])
def test_slab_schedule_is_exact_under_random_interleaving(world, tiles, stages, axis, seed):
  kw = random_problem(domain=(9, 12, 8), axis=axis, tt=11, seed=seed, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  out = SlabEmulator(kw, world, tiles, stages, max_lead=4, need_rule=3, seed=seed).run()
  assert np.isfinite(out).all()
  np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("world,tiles,stages,axis,seed", [(2, 1, 4, 0, 0), (2, 2, 3, 1, 1), (1, 1, 5, 2, 2),
                                                        (4, 1, 2, 2, 3)])
def test_l2_discard_never_hits_a_line_a_courier_or_a_ghost_reader_needs(world, tiles, stages, axis, seed):
  """With the consumed lines of a tile's exclusive columns poisoned (discard.global.L2), a slab
  run is still exact: couriers read edge columns only, and those are never discarded."""
  kw = random_problem(domain=(9, 16, 8), axis=axis, tt=11, seed=seed, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  out = SlabEmulator(kw, world, tiles, stages, max_lead=4, seed=seed, discard=True).run()
  assert np.isfinite(out).all()
  np.testing.assert_array_equal(out, ref)


def test_discarding_the_edge_columns_is_detected():
  """Negative control: were the tile-edge columns discarded too, a neighbour tile's halo load or a
  courier would pick up a poisoned line."""
  kw = random_problem(domain=(9, 16, 8), axis=0, tt=11, seed=0, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  wrong = 0
  for seed in range(3):
    out = SlabEmulator(kw, 2, 1, 4, max_lead=4, seed=seed, discard=True,
                       discard_edge_columns=True).run()
    wrong += not np.array_equal(out, ref)
  assert wrong >= 1


def test_random_geometries():
  """Forty random (ranks, tiles per rank, tile width, stages, throttle, orientation, length, discard)
  draws; a 25-minute run of the same loop (11 800 draws, throttle >= 2) found no mismatch and no
  deadlock, and a throttle of 1 plane deadlocks as section 4.5 of DESIGN.md says it must."""
  rng = np.random.default_rng(7)
  for _ in range(40):
    world, tiles, per = (int(rng.integers(1, 5)), int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    X, Z, tt = int(rng.integers(3, 10)), int(rng.choice([6, 8])), int(rng.integers(2, 14))
    stages, axis, seed = int(rng.integers(1, 9)), int(rng.integers(0, 3)), int(rng.integers(0, 10**6))
    lead, discard = int(rng.integers(2, 8)), bool(rng.integers(0, 2))
    kw = random_problem(domain=(X, world * tiles * per, Z), axis=axis, tt=tt, seed=seed,
                        output_steps=(max(0, tt - 5), tt, 2))
    out = SlabEmulator(kw, world, tiles, stages, max_lead=lead, seed=seed, discard=discard).run()
    np.testing.assert_array_equal(out, fdtd_numpy.fdtdz(**kw))


def test_couriers_may_lag_arbitrarily():
  """A courier that is scheduled rarely (here: 20x less often than a tile) delays its neighbour
  but never corrupts it."""
  kw = random_problem(domain=(9, 12, 8), axis=0, tt=11, seed=3, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  emu = SlabEmulator(kw, 2, 2, 4, max_lead=4, seed=3)
  rng = np.random.default_rng(3)
  keys = list(emu.actor.keys())
  while True:
    live = [key for key in keys if emu.actor[key]["n"] < emu.tt]
    tiles = [key for key in live if emu.ready(*key)]
    cours = [c for c in emu.couriers if emu.courier_ready(*c)]
    if not live and not cours:
      break
    assert tiles or cours, "deadlock"
    if tiles and (not cours or rng.random() < 0.95):
      emu.advance(*tiles[rng.integers(len(tiles))])
    else:
      emu.courier_advance(*cours[rng.integers(len(cours))])
  np.testing.assert_array_equal(emu.out, ref)


def test_courier_write_after_read_rule_is_implied_by_the_chain():
  """The kernel also makes an edge tile wait, before it overwrites a plane of its edge column,
  until the courier has carried off what the step two back wrote there (``ctl.cour``).  The
  emulator shows that wait never binds: step n at plane k needs the neighbour's step n-1 at k+3
  (mirror slot), which needed this rank's step n-2 at k+6 -- forwarded by the courier only after
  the copy.  So the result is exact without the rule, and with it the rule is never the one that
  holds a tile back.  It stays in the kernel as a guard that costs one shared-memory read."""
  kw = random_problem(domain=(9, 12, 8), axis=0, tt=11, seed=0, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  for seed in range(4):
    emu = SlabEmulator(kw, 2, 1, 4, max_lead=4, seed=seed, courier_war_rule=False)
    np.testing.assert_array_equal(emu.run(), ref)
    assert emu.war_would_block == 0


def test_counter_forwarded_before_the_data_is_detected():
  """Negative control: a courier that forwards the counter first and copies later lets the
  neighbour read ghost planes that have not arrived."""
  kw = random_problem(domain=(9, 12, 8), axis=0, tt=11, seed=0, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  wrong = 0
  for seed in range(4):
    out = SlabEmulator(kw, 2, 1, 4, max_lead=4, seed=seed, counter_before_data=True).run()
    wrong += not np.array_equal(out, ref, equal_nan=True)
  assert wrong >= 1


def test_weaker_lookahead_across_the_slab_edge_is_detected():
  """Negative control: the k+3 rule weakened to k+1 must break exactness across slab edges too."""
  kw = random_problem(domain=(9, 12, 8), axis=0, tt=11, seed=0, output_steps=(3, 11, 2))
  ref = fdtd_numpy.fdtdz(**kw)
  wrong = 0
  for seed in range(4):
    out = SlabEmulator(kw, 2, 1, 4, max_lead=4, need_rule=1, seed=seed).run()
    wrong += not np.array_equal(out, ref, equal_nan=True)
  assert wrong >= 1
