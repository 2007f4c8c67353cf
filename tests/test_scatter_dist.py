"""Port-batch sharding of scatter() across ranks (first axis): world_size-2
gloo run on the CPU with the oracle engine must reproduce the single-process S-matrix."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
  from pjz_b200 import SimParams, mode
  omega = np.array([2 * np.pi / 37])
  eps = np.ones((3, 20, 30, 20), np.float32)
  eps[:, :, 9:21, 8:12] = 12.25
  beta, exc, _, _ = mode(eps[:, 3:4], omega, 2)
  modes = [exc[..., 0], exc[..., 1], exc[..., 0]]
  betas = [beta[:, 0], beta[:, 1], beta[:, 0]]
  p = SimParams(omega_range=(omega[0], omega[0]), tt=200, dt=0.5, absorption_padding=3,
                absorption_coeff=4e-4, pml_widths=(4, 4), use_reduced_precision=False,
                domain_zz=28)
  return eps, omega, modes, betas, (3, 4, 16), (True, True, False), p


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from oracle import fdtd_c
  from pjz_b200 import scatter
  eps, omega, modes, betas, pos, fwd, p = _problem()
  sv = scatter(eps, omega, modes, betas, pos, fwd, p, engine=fdtd_c.fdtdz)
  if rank == 0:
    torch.save([[s.clone() for s in row] for row in sv], out)
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_scatter_matches_single_process(tmp_path):
  from oracle import fdtd_c
  from pjz_b200 import scatter
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "sv.pt")
  mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
  got = torch.load(out)
  eps, omega, modes, betas, pos, fwd, p = _problem()
  want = scatter(eps, omega, modes, betas, pos, fwd, p, engine=fdtd_c.fdtdz)
  for i in range(3):
    for j in range(3):
      torch.testing.assert_close(got[i][j], want[i][j], rtol=0, atol=0)


def _worker_decomposed(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from pjz_b200 import decomposed_engine, scatter
  from tests.test_decomp import OracleSlabY
  eps, omega, modes, betas, pos, fwd, p = _problem()
  p = p._replace(tt=60)
  engine = decomposed_engine("y", ghost=3, make_slab=OracleSlabY)
  sv = scatter(eps, omega, modes[:2], betas[:2], pos[:2], fwd[:2], p, engine=engine)
  torch.save([[s.clone() for s in row] for row in sv], out + f".{rank}")
  dist.barrier()
  dist.destroy_process_group()


def test_scatter_over_a_domain_decomposed_engine(tmp_path):
  """The second axis of SURVEY.md 8(e) through the reference-facing call: every port is solved by
  BOTH ranks (y-slabs, ghost zones over gloo, oracle-backed slabs), every rank gets the whole
  S-matrix, and it equals the single-domain one exactly."""
  from oracle import fdtd_numpy
  from pjz_b200 import scatter
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "sv_dd.pt")
  mp.spawn(_worker_decomposed, args=(2, port, out), nprocs=2, join=True)
  eps, omega, modes, betas, pos, fwd, p = _problem()
  p = p._replace(tt=60)
  want = scatter(eps, omega, modes[:2], betas[:2], pos[:2], fwd[:2], p, engine=fdtd_numpy.fdtdz)
  for rank in range(2):
    got = torch.load(out + f".{rank}")
    for i in range(2):
      for j in range(2):
        torch.testing.assert_close(got[i][j], want[i][j], rtol=0, atol=0)


def _gpu_worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from pjz_b200 import scatter
  eps, omega, modes, betas, pos, fwd, p = _problem()
  e = torch.from_numpy(eps).cuda().requires_grad_(True)
  sv = scatter(e, omega, modes, betas, pos, fwd, p)          # CUDA engine, ports dealt to the ranks
  loss = sum((s.abs() ** 2).sum() for row in sv for s in row)
  loss.backward()                                            # fused product-reduce on every rank
  if rank == 0:
    torch.save({"sv": [[s.detach().cpu() for s in row] for row in sv], "grad": e.grad.cpu()}, out)
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_scatter_matches_single_gpu(built, tmp_path):
  """Port batch over NCCL: 3 ports on 2 GPUs (engine runs sharded, phasor fields broadcast),
  S-matrix and gradient equal to the single-GPU call."""
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  from pjz_b200 import scatter
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  out = str(tmp_path / "sv_gpu.pt")
  mp.spawn(_gpu_worker, args=(2, port, out), nprocs=2, join=True)
  got = torch.load(out)
  eps, omega, modes, betas, pos, fwd, p = _problem()
  e = torch.from_numpy(eps).cuda().requires_grad_(True)
  want = scatter(e, omega, modes, betas, pos, fwd, p)
  sum((s.abs() ** 2).sum() for row in want for s in row).backward()
  for i in range(3):
    for j in range(3):
      torch.testing.assert_close(got["sv"][i][j], want[i][j].detach().cpu(), rtol=1e-6, atol=1e-9)
  torch.testing.assert_close(got["grad"], e.grad.cpu(), rtol=1e-5, atol=1e-9)
