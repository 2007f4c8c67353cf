"""x-slab domain decomposition (pjz_b200/_decomp.py).

CPU: the orchestration (slab inputs, ghost planes, halo exchange order, source ownership,
snapshot cropping) is driven with an ORACLE-backed slab engine over gloo, world sizes 1-3, and
must reproduce the single-domain oracle run EXACTLY.  GPU: the same driver over the C ABI's
stepping session must be bit-identical to the one-call engine (1 GPU, self-wrapped ghosts; and
2 ranks when two GPUs are present).
"""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSlab:
  """Slab engine backed by the float64 NumPy spec (tests only)."""

  def __init__(self, loc):
    from oracle import fdtd_numpy as spec
    self.spec = spec
    self.loc = loc
    eps = np.asarray(loc["epsilon"], np.float32)
    self.st = spec.State(eps, loc["dt"], loc["absorption_mask"], loc["pml_kappa"],
                         loc["pml_sigma"], loc["pml_alpha"], loc["pml_widths"], loc["offset"],
                         np.float64)
    self.sf = np.asarray(loc["source_field"], np.float64)
    self.wf = np.asarray(loc["source_waveform"], np.float32).astype(np.float64)
    self.axis = spec.source_axis(self.sf)
    self.outs = list(range(*loc["output_steps"]))
    _, xx, yy, zz = eps.shape
    self.sub = (xx, yy, zz)
    self.out = np.zeros((len(self.outs), 3, xx, yy, zz), np.float32)
    self.E = torch.from_numpy(self.st.E)      # shared memory views (3, X, Y, Z)
    self.H = torch.from_numpy(self.st.H)

  def step_h(self):
    self.st.step_h()

  def step_e(self, n):
    self.st.step_e()
    self.st.add_source(self.sf, self.wf[n], self.loc["source_position"], self.axis)
    if n in self.outs:
      ox, oy, oz = self.loc["offset"]
      xx, yy, zz = self.sub
      self.out[self.outs.index(n)] = self.st.E[:, ox:ox + xx, oy:oy + yy, oz:oz + zz]

  def snapshots(self):
    return self.out


class OracleSlabY(OracleSlab):
  """y-slab engine (whole-step ``advance``) backed by the float64 NumPy spec (tests only)."""

  def __init__(self, loc):
    super().__init__(loc)
    self._state = [self.E, self.H, torch.from_numpy(self.st.psiH), torch.from_numpy(self.st.psiE)]

  def state(self, n):
    return self._state

  def advance(self, n0, nsteps):
    for n in range(n0, n0 + nsteps):
      self.step_h()
      self.step_e(n)


def _problems():
  from tests.problems import random_problem
  return [
      random_problem(domain=(12, 10, 8), axis=0, pml=(2, 3), tt=14, seed=21, output_steps=(5, 14, 4),
                     src_pos=4),
      random_problem(domain=(11, 9, 8), axis=0, pml=(2, 2), tt=12, seed=22, output_steps=(3, 12, 3),
                     src_pos=6),                 # plane 6 / 5 straddle the 2-rank cut (5|6)
      random_problem(domain=(10, 8, 8), axis=1, pml=(0, 3), tt=12, seed=23, output_steps=(0, 12, 5)),
      random_problem(domain=(9, 10, 12), axis=2, pml=(3, 3), tt=12, seed=24, output_steps=(11, 12, 1)),
  ]


def _worker(rank, world, port, out_path):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from pjz_b200._decomp import fdtdz_decomposed
  outs = [fdtdz_decomposed(**kw, make_slab=OracleSlab).numpy() for kw in _problems()]
  if rank == 0:
    np.savez(out_path, *outs)
  dist.barrier()
  dist.destroy_process_group()


def _problems_y():
  from tests.problems import random_problem
  return [
      random_problem(domain=(9, 24, 8), axis=0, pml=(2, 3), tt=14, seed=31, output_steps=(5, 14, 4)),
      random_problem(domain=(8, 21, 8), axis=1, pml=(2, 2), tt=13, seed=32, output_steps=(3, 13, 3),
                     src_pos=10),                # y-plane 10 / 9 straddle the 2-rank cut (10|11)
      random_problem(domain=(7, 18, 12), axis=2, pml=(3, 3), tt=12, seed=33, output_steps=(11, 12, 1)),
      random_problem(domain=(6, 20, 8), axis=1, pml=(0, 3), tt=9, seed=34, output_steps=(0, 9, 4),
                     src_pos=0),
  ]


def _worker_y(rank, world, port, out_path, ghost):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from pjz_b200._decomp import fdtdz_decomposed_y
  outs = [fdtdz_decomposed_y(**kw, ghost=ghost, make_slab=OracleSlabY).numpy()
          for kw in _problems_y()]
  if rank == 0:
    np.savez(out_path, *outs)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize("world,ghost", [(2, 3), (3, 2), (2, 5)])
def test_y_decomposed_oracle_matches_single_domain(world, ghost, tmp_path):
  """Ghost-zone scheme: `ghost` steps per exchange, E/H/psi columns, all source orientations."""
  from oracle import fdtd_numpy
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "ddy.npz")
  mp.spawn(_worker_y, args=(world, port, out, ghost), nprocs=world, join=True)
  got = np.load(out)
  for i, kw in enumerate(_problems_y()):
    np.testing.assert_array_equal(got[f"arr_{i}"], fdtd_numpy.fdtdz(**kw))


def test_y_single_rank_wraps_onto_itself():
  from oracle import fdtd_numpy
  from pjz_b200._decomp import fdtdz_decomposed_y
  for kw in _problems_y()[:3]:
    for ghost in (1, 4):
      out = fdtdz_decomposed_y(**kw, ghost=ghost, make_slab=OracleSlabY).numpy()
      np.testing.assert_array_equal(out, fdtd_numpy.fdtdz(**kw))
  with pytest.raises(NotImplementedError):       # y-plane source at column 0: owned + ghost image
    fdtdz_decomposed_y(**_problems_y()[3], ghost=2, make_slab=OracleSlabY)


def test_y_slab_inputs():
  from pjz_b200._decomp import local_problem_y
  from tests.problems import random_problem
  kw = random_problem(domain=(6, 12, 8), sub=(4, 5, 3), offset=(1, 4, 2), axis=1, tt=6, seed=1,
                      src_pos=7)
  loc, nloc, crop = local_problem_y(kw, 1, 2, 2)   # owns columns 6..11, ghosts 4,5 and 0,1
  assert nloc == 6 and loc["epsilon"].shape == (3, 4, 10, 3) and loc["offset"] == (1, 0, 2)
  assert loc["absorption_mask"].shape == (3, 6, 10)
  np.testing.assert_array_equal(loc["absorption_mask"][:, :, 0], kw["absorption_mask"][:, :, 4])
  np.testing.assert_array_equal(loc["absorption_mask"][:, :, 9], kw["absorption_mask"][:, :, 1])
  # epsilon columns: global 4..8 are the sub-volume; beyond it the edge is replicated
  np.testing.assert_array_equal(loc["epsilon"][:, :, 0], kw["epsilon"][:, :, 0])
  np.testing.assert_array_equal(loc["epsilon"][:, :, 9], kw["epsilon"][:, :, 0])   # global col 1
  np.testing.assert_array_equal(loc["epsilon"][:, :, 5], kw["epsilon"][:, :, 4])   # global col 9
  assert loc["source_position"] == 3 and loc["source_waveform"].any()              # global 7
  assert crop == (2, 5, 2, 5)                      # global 6..8 -> local 2..4, out 2..4
  loc0, _, crop0 = local_problem_y(kw, 0, 2, 2)    # owns 0..5; ghosts 10,11 and 6,7: source in ghost
  assert loc0["source_position"] == 9 and loc0["source_waveform"].any()
  assert crop0 == (6, 8, 0, 2)
  with pytest.raises(ValueError):
    local_problem_y(kw, 0, 4, 4)


@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_oracle_matches_single_domain(world, tmp_path):
  from oracle import fdtd_numpy
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "dd.npz")
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  got = np.load(out)
  for i, kw in enumerate(_problems()):
    np.testing.assert_array_equal(got[f"arr_{i}"], fdtd_numpy.fdtdz(**kw))


def test_single_rank_wraps_onto_itself():
  from oracle import fdtd_numpy
  from pjz_b200._decomp import fdtdz_decomposed
  for kw in _problems():
    out = fdtdz_decomposed(**kw, make_slab=OracleSlab).numpy()
    np.testing.assert_array_equal(out, fdtd_numpy.fdtdz(**kw))


def test_slab_inputs():
  from pjz_b200._decomp import local_problem, slab_bounds
  from tests.problems import random_problem
  kw = random_problem(domain=(11, 9, 8), sub=(5, 4, 3), offset=(4, 2, 1), axis=0, tt=6, seed=1,
                      src_pos=6)
  assert [slab_bounds(11, 3, r) for r in range(3)] == [(0, 3), (3, 7), (7, 11)]
  loc, nloc, crop = local_problem(kw, 1, 3)       # owns planes 3..6, ghosts 2 and 7
  assert nloc == 4 and loc["epsilon"].shape == (3, 6, 4, 3) and loc["offset"] == (0, 2, 1)
  assert loc["absorption_mask"].shape == (3, 6, 9)
  # epsilon is edge-replicated in x: local planes 0..2 (global 2,3,4) all see sub-volume plane 0
  np.testing.assert_array_equal(loc["epsilon"][:, 0], kw["epsilon"][:, 0])
  np.testing.assert_array_equal(loc["epsilon"][:, 2], kw["epsilon"][:, 0])
  np.testing.assert_array_equal(loc["epsilon"][:, 4], kw["epsilon"][:, 2])
  # source plane 6 is local plane 4 (owned); plane 5 (channel 1) is local 3 (owned)
  assert loc["source_position"] == 4 and loc["source_waveform"].any()
  assert crop == (2, 5, 0, 3)                     # global planes 4..6 -> local 2..4, out 0..2
  loc2, nloc2, crop2 = local_problem(kw, 2, 3)    # owns 7..10: plane 6 is its low ghost
  assert not loc2["source_waveform"].any()
  assert crop2 == (1, 3, 3, 5)                    # global 7..8 -> local 1..2, out 3..4
  with pytest.raises(ValueError):
    local_problem(kw, 0, 12)


def test_x_source_on_plane_zero_of_a_slab_that_owns_the_wrap():
  """with one rank the slab owns plane 0 AND plane X-1; ownership of channel 1's plane
  (p-1) % X must come from the global index.  An active second channel there cannot be injected
  by one local plane pair and must raise instead of vanishing silently; a silent channel works."""
  from oracle import fdtd_numpy
  from pjz_b200._decomp import fdtdz_decomposed, local_problem
  from tests.problems import random_problem
  kw = random_problem(domain=(9, 7, 8), axis=0, tt=9, seed=5, src_pos=0, output_steps=(3, 9, 2))
  with pytest.raises(NotImplementedError):
    local_problem(kw, 0, 1)
  quiet = dict(kw)
  quiet["source_waveform"] = kw["source_waveform"].copy()
  quiet["source_waveform"][:, 1] = 0
  loc, nloc, _ = local_problem(quiet, 0, 1)
  assert loc["source_position"] == 1 and loc["source_waveform"][:, 0].any()
  out = fdtdz_decomposed(**quiet, make_slab=OracleSlab).numpy()
  np.testing.assert_array_equal(out, fdtd_numpy.fdtdz(**quiet))
  # two ranks: channel 1's plane X-1 belongs to the LAST rank, channel 0's plane 0 to the first
  a, _, _ = local_problem(kw, 0, 2)
  b, nb, _ = local_problem(kw, 1, 2)
  assert a["source_position"] == 1 and a["source_waveform"][:, 0].any() and not a["source_waveform"][:, 1].any()
  assert b["source_position"] == nb + 1 and b["source_waveform"][:, 1].any() and not b["source_waveform"][:, 0].any()


# ---- GPU ------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_session_driver_equals_one_call_engine(built):
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import fdtdz_decomposed
  from tests.problems import random_problem
  for axis, reduced in [(0, False), (1, False), (2, False), (0, True)]:
    kw = random_problem(domain=(20, 18, 40), axis=axis, pml=(5, 6), tt=30, seed=60 + axis,
                        output_steps=(10, 30, 7), reduced=reduced, src_pos=8 if axis == 0 else None)
    dev = dict(kw)
    dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
    dev["launch_params"] = {"kernel": "twopass"}
    want = fdtdz_jax.fdtdz(**dev).cpu().numpy()
    got = fdtdz_decomposed(**kw).cpu().numpy()   # world = 1: the slab wraps onto itself
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_y_session_driver_equals_one_call_engine(built):
  """World 1 (the slab wraps onto itself through its own ghosts): the systolic stepping session
  (lean geometry) and the per-step session (any geometry) both reproduce the one-call engine."""
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import YSlabRun, fdtdz_decomposed_y
  from tests.problems import random_problem
  cases = [((10, 40, 128), (16, 16), 0, 6, False), ((9, 36, 128), (4, 6), 1, 4, False),
           ((8, 30, 126), (0, 0), 2, 7, False), ((12, 20, 40), (5, 6), 0, 3, False),
           ((10, 24, 32), (4, 4), 1, 2, True),
           # short columns: the sub-warp lean kernel advances the slab (fp16 and fp32 storage)
           ((10, 40, 128), (16, 16), 0, 6, True), ((9, 36, 64), (4, 6), 1, 4, False),
           ((8, 30, 60), (0, 0), 2, 7, True), ((11, 50, 100), (12, 9), 1, 5, True)]
  for domain, pml, axis, ghost, reduced in cases:
    kw = random_problem(domain=domain, axis=axis, pml=pml, tt=27, seed=80 + axis,
                        output_steps=(9, 27, 5), reduced=reduced)
    dev = dict(kw)
    dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
    want = fdtdz_jax.fdtdz(**dev).cpu().numpy()
    got = fdtdz_decomposed_y(**kw, ghost=ghost).cpu().numpy()
    np.testing.assert_array_equal(got, want)
  run = YSlabRun(random_problem(domain=(10, 40, 128), pml=(16, 16), tt=8, seed=1), ghost=4)
  assert run.slab.kernel == "systolic_lean" and run.slab.pingpong
  run.close()
  run = YSlabRun(random_problem(domain=(10, 40, 128), pml=(16, 16), tt=8, seed=1, reduced=True), ghost=4)
  assert run.slab.kernel == "systolic_lean" and run.slab.pingpong
  run.close()


def _gpu_worker_y(rank, world, port, out_path):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from pjz_b200._decomp import fdtdz_decomposed_y
  from tests.problems import random_problem
  kw = random_problem(domain=(24, 64, 128), axis=1, pml=(16, 16), tt=40, seed=78,
                      output_steps=(20, 40, 6), src_pos=31)
  out = fdtdz_decomposed_y(**kw, ghost=6).cpu().numpy()
  if rank == 0:
    np.save(out_path, out)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_y_decomposition_is_bit_exact(built, tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  from oracle import fdtd_c
  from tests.problems import random_problem
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "ddy.npy")
  mp.spawn(_gpu_worker_y, args=(2, port, out), nprocs=2, join=True)
  kw = random_problem(domain=(24, 64, 128), axis=1, pml=(16, 16), tt=40, seed=78,
                      output_steps=(20, 40, 6), src_pos=31)
  np.testing.assert_array_equal(np.load(out), fdtd_c.fdtdz(**kw))


def _gpu_worker(rank, world, port, out_path):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from pjz_b200._decomp import fdtdz_decomposed
  from tests.problems import random_problem
  kw = random_problem(domain=(40, 24, 32), axis=0, pml=(4, 6), tt=40, seed=77,
                      output_steps=(20, 40, 6), src_pos=20)
  out = fdtdz_decomposed(**kw).cpu().numpy()
  if rank == 0:
    np.save(out_path, out)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_decomposition_is_bit_exact(built, tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  from oracle import fdtd_c
  from tests.problems import random_problem
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "dd.npy")
  mp.spawn(_gpu_worker, args=(2, port, out), nprocs=2, join=True)
  kw = random_problem(domain=(40, 24, 32), axis=0, pml=(4, 6), tt=40, seed=77,
                      output_steps=(20, 40, 6), src_pos=20)
  np.testing.assert_array_equal(np.load(out), fdtd_c.fdtdz(**kw))


# ---- in-kernel halo exchange (peer-mapped stores, one launch per GPU) ---------------------------------

P2P_CASES = [
    # domain, pml, axis, src_pos, steps
    ((10, 1, 128), (16, 16), 0, 4, 27),      # one owned column: the edge column is first AND last
    ((9, 2, 128), (4, 6), 2, 20, 27),
    ((9, 14, 125), (3, 5), 0, 3, 27),        # even tile width: the last owned column is a column A
    ((11, 15, 128), (16, 16), 1, 7, 45),     # one full tile (odd width: last owned column is a B)
    ((12, 16, 128), (16, 16), 2, 40, 33),    # two tiles
    ((8, 47, 126), (0, 0), 1, 46, 27),       # y source on the last column: its partner is column 45
    ((16, 64, 128), (16, 16), 0, 5, 64),     # several tiles and pipeline rounds
]


@pytest.mark.gpu
@pytest.mark.parametrize("domain,pml,axis,src_pos,tt", P2P_CASES)
def test_p2p_slab_wrapping_onto_itself_equals_one_call_engine(built, domain, pml, axis, src_pos, tt):
  """World 1: the slab's low and high neighbour are the slab itself, so the couriers store into
  the slab's OWN ghost columns and the edge tiles watch their own mirror slots -- the same code
  path as between two GPUs, minus NVLink.  Must reproduce the periodic one-call engine bit for
  bit (and therefore the C oracle)."""
  from oracle import fdtd_c
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import fdtdz_decomposed_p2p
  from tests.problems import random_problem
  kw = random_problem(domain=domain, axis=axis, pml=pml, tt=tt, seed=90 + axis, src_pos=src_pos,
                      output_steps=(tt // 3, tt, 5), absorb_pad=min(3, domain[1] // 2))
  got = fdtdz_decomposed_p2p(**kw).cpu().numpy()
  np.testing.assert_array_equal(got, fdtd_c.fdtdz(**kw))
  dev = dict(kw)
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  np.testing.assert_array_equal(got, fdtdz_jax.fdtdz(**dev).cpu().numpy())


@pytest.mark.gpu
def test_p2p_slab_tilings_and_long_run(built):
  """Explicit tilings (tile width / stage count through launch_params) and a longer run on a
  wider slab: every plan gives the bits of the one-call engine."""
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import fdtdz_decomposed_p2p
  from tests.problems import random_problem
  kw = random_problem(domain=(24, 96, 128), axis=2, pml=(16, 16), tt=150, seed=5,
                      output_steps=(60, 150, 41), absorb_pad=8, absorb_coeff=1e-3)
  dev = dict(kw)
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  want = fdtdz_jax.fdtdz(**dev).cpu().numpy()
  for lp in (None, {"tile_y": 5, "stages": 3}, {"tile_y": 13, "stages": 2}, {"tile_y": 1, "stages": 1}):
    got = fdtdz_decomposed_p2p(**{**kw, "launch_params": lp}).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
def test_p2p_slab_chained_launches_and_the_reset_guard(built):
  """A slab session advances in several launches (the ghosts hold the neighbour's last state), and
  a launch without b200fdtd_session_slab_reset since the previous one is refused: the counters the
  previous launch left behind would make every dependency look met."""
  from pjz_b200 import fdtdz_jax
  from pjz_b200._decomp import P2PSlabRun
  from tests.problems import random_problem
  kw = random_problem(domain=(16, 40, 128), axis=0, pml=(16, 16), tt=61, seed=17, src_pos=5,
                      output_steps=(9, 61, 7), absorb_pad=4)
  dev = dict(kw)
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  want = fdtdz_jax.fdtdz(**dev)
  run = P2PSlabRun(kw)
  slab = run.slab
  slab.advance(0, 23)
  torch.cuda.synchronize()
  rc = slab.L.b200fdtd_session_advance(slab.session, 23, 10, slab._stream())
  assert rc != 0 and "slab_reset" in fdtdz_jax._last_error()
  slab.advance(23, 1)                    # odd and even starting steps, one-step launches
  slab.advance(24, 12)
  slab.advance(36, 25)
  got = run.gathered_snapshots()
  run.close()
  assert torch.equal(got, want)


@pytest.mark.gpu
def test_field_through_the_decomposed_engine(built):
  """``pjz_b200.field(engine=decomposed_engine("p2p"))``: the reference-facing call over the
  y-slab sessions (one rank: the slab wraps onto itself) gives the phasors of the default engine."""
  from pjz_b200 import SimParams, decomposed_engine, field
  rng = np.random.default_rng(3)
  eps = np.ones((3, 24, 32, 20), np.float32)
  eps[:, :, 10:22, 8:12] = 12.25
  src = rng.standard_normal((2, 1, 32, 20)).astype(np.float32)
  omega = np.array([2 * np.pi / 37])
  p = SimParams(omega_range=(omega[0], omega[0]), tt=120, dt=0.5, absorption_padding=3,
                absorption_coeff=4e-4, pml_widths=(16, 16), use_reduced_precision=False,
                domain_zz=128)
  e = torch.from_numpy(eps).cuda()
  want = field(e, src, omega, 4, p)
  got = field(e, src, omega, 4, p, engine=decomposed_engine("p2p"))
  assert bool(torch.isfinite(want).all()) and float(want.abs().max()) > 0
  assert torch.equal(got, want)
  with pytest.raises(NotImplementedError):
    field(e, src, omega, 4, p, engine=decomposed_engine("p2p"), fuse_projection=True)


def _gpu_worker_p2p(rank, world, port, out_path):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from pjz_b200._decomp import fdtdz_decomposed_p2p
  from tests.problems import random_problem
  outs = []
  for axis, src_pos in ((1, 31), (1, 32), (0, 7), (2, 50)):   # y sources next to / on the cut
    kw = random_problem(domain=(24, 64, 128), axis=axis, pml=(16, 16), tt=80, seed=78 + axis,
                        output_steps=(20, 80, 12), src_pos=src_pos)
    outs.append(fdtdz_decomposed_p2p(**kw).cpu().numpy())
  if rank == 0:
    np.savez(out_path, *outs)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_p2p_decomposition_is_bit_exact(built, tmp_path):
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  from oracle import fdtd_c
  from tests.problems import random_problem
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "p2p.npz")
  mp.spawn(_gpu_worker_p2p, args=(2, port, out), nprocs=2, join=True)
  got = np.load(out)
  for i, (axis, src_pos) in enumerate(((1, 31), (1, 32), (0, 7), (2, 50))):
    kw = random_problem(domain=(24, 64, 128), axis=axis, pml=(16, 16), tt=80, seed=78 + axis,
                        output_steps=(20, 80, 12), src_pos=src_pos)
    np.testing.assert_array_equal(got[f"arr_{i}"], fdtd_c.fdtdz(**kw))
