"""x-slab domain decomposition of one engine call across the GPUs of a box.

Second multi-GPU axis of SURVEY.md 8(e) (the first, the port batch, is in ``_field.py``):
a domain too large -- or too slow -- for one GPU is cut into contiguous x-slabs, one rank
(process, GPU) per slab.  Every rank holds its ``nloc`` planes plus one ghost plane on each
side and advances half-steps with the per-step kernels through the session interface of the
C ABI (``b200fdtd_session_*``, include/b200fdtd.h).  After every half-step the faces are
exchanged, ring-wrapped for the periodic x boundary:

    after the H half-step:  H[last owned plane]  -> right neighbour's low  ghost plane
    after the E half-step:  E[first owned plane] -> left  neighbour's high ghost plane

(the H update reads E at x+1, the E update reads H at x-1).  Transport is
``torch.distributed`` point-to-point (NCCL over NVLink on the GPUs, gloo in the CPU tests);
a face is 3*Y*Z words, ~0.1 % of a slab's per-step traffic for the BASELINE metalens
configuration.  The ghost planes are also swept by the kernels (with garbage neighbours) and
overwritten by the next exchange before anything reads them, so every owned cell sees exactly
the inputs it sees in the single-GPU run: the result is bit-identical.

The reference has no counterpart (fdtd-z is single-GPU); the interface is the engine call's own
keyword list plus ``group``.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(X, world, rank):
  return rank * X // world, (rank + 1) * X // world


def _np(a):
  if isinstance(a, torch.Tensor):
    return a.detach().cpu().numpy()
  return np.asarray(a)


def local_problem(kw, rank, world):
  """Engine kwargs of this rank's slab (with ghost planes) + how to crop its snapshots.

  The local domain is ``nloc + 2`` planes: local plane 0 / nloc+1 are the ghosts (global planes
  x0-1 / x1, periodic).  epsilon is extended in x on the host (edge replication, as the engine
  would do) so that the slab -- ghosts included -- gets its own permittivity; y/z keep the
  engine's own edge replication through ``offset``.
  """
  eps = _np(kw["epsilon"]).astype(np.float32)
  mask = _np(kw["absorption_mask"]).astype(np.float32)
  sf = _np(kw["source_field"]).astype(np.float32)
  X = mask.shape[1]
  ox, oy, oz = (int(o) for o in kw["offset"])
  xx = eps.shape[1]
  x0, x1 = slab_bounds(X, world, rank)
  nloc = x1 - x0
  if nloc < 1:
    raise ValueError(f"more ranks ({world}) than x-planes ({X})")
  planes = np.arange(x0 - 1, x1 + 1) % X                     # global plane of each local plane
  eps_x = np.clip(planes - ox, 0, xx - 1)                    # edge replication in x
  loc = dict(kw)
  loc["epsilon"] = np.ascontiguousarray(eps[:, eps_x])
  loc["absorption_mask"] = np.ascontiguousarray(mask[:, planes])
  loc["offset"] = (0, oy, oz)
  axis = 2 if sf.ndim == 5 else (0 if sf.shape[1] == 1 else 1)
  p = int(kw["source_position"])
  wf = _np(kw["source_waveform"]).astype(np.float32)
  if axis == 0:
    # channel 0 acts on global plane p, channel 1 on p-1; a rank keeps a channel only if the
    # plane it acts on is one of its OWNED planes (a hit on a ghost plane would be harmless --
    # the ghost is overwritten by the next exchange -- but masking keeps the intent explicit).
    lp = (p - (x0 - 1)) % X                                  # local index of global plane p
    lp1 = (p - 1 - (x0 - 1)) % X                             # ... and of plane p-1 (periodic)
    wf = wf.copy()
    own0 = 1 <= lp <= nloc
    own1 = 1 <= lp1 <= nloc
    if own1 and lp1 != lp - 1 and wf[:, 1].any():
      # plane p-1 wraps around the periodic boundary onto a non-adjacent local plane
      raise NotImplementedError("x source at plane 0 with an active second channel")
    if not own0:
      wf[:, 0] = 0
    if not own1:
      wf[:, 1] = 0
    if not (own0 or own1) or lp > nloc + 1:
      lp = 1
      wf[:] = 0
    loc["source_position"] = int(lp)
    loc["source_waveform"] = wf
    loc["source_field"] = sf
  elif axis == 1:
    loc["source_field"] = np.ascontiguousarray(sf[:, planes])
  else:
    loc["source_field"] = np.ascontiguousarray(sf[:, :, planes])
  # snapshot crop: owned planes that fall inside the caller's sub-volume [ox, ox+xx)
  g0, g1 = max(ox, x0), min(ox + xx, x1)
  crop = None
  if g1 > g0:
    crop = (g0 - x0 + 1, g1 - x0 + 1, g0 - ox, g1 - ox)      # (local lo, local hi, out lo, out hi)
  return loc, nloc, crop


class CudaSlab:
  """Slab engine over the C ABI's stepping session (per-step CUDA kernels, in place)."""

  def __init__(self, loc, device):
    from . import fdtdz_jax as shim
    self.shim = shim
    L = shim.lib()
    loc = dict(loc)
    loc["launch_params"] = {"kernel": "twopass"}
    self.d = shim.make_desc(**{k: loc[k] for k in (
        "epsilon", "dt", "source_field", "source_waveform", "source_position", "absorption_mask",
        "pml_kappa", "pml_sigma", "pml_alpha", "pml_widths", "output_steps",
        "use_reduced_precision", "launch_params", "offset")})
    self.device = torch.device(device)
    names = ("epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa",
             "pml_sigma", "pml_alpha")
    self.inputs = [torch.from_numpy(np.ascontiguousarray(_np(loc[k]), np.float32)).to(self.device)
                   for k in names]
    L.b200fdtd_session_workspace_bytes.restype = ctypes.c_size_t
    nbytes = L.b200fdtd_session_workspace_bytes(ctypes.byref(self.d))
    if nbytes == 0:
      raise RuntimeError(shim._last_error())
    self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
    nout = L.b200fdtd_num_outputs(ctypes.byref(self.d))
    self.out = torch.zeros((nout, 3, self.d.xx, self.d.yy, self.d.zz), dtype=torch.float32,
                           device=self.device)
    self.session = ctypes.c_void_p()
    L.b200fdtd_session_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_void_p]
    L.b200fdtd_session_step_h.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_session_step_e.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.b200fdtd_session_layout.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_session_destroy.argtypes = [ctypes.c_void_p]
    L.b200fdtd_session_destroy.restype = None
    self.L = L
    with torch.cuda.device(self.device):
      rc = L.b200fdtd_session_create(
          ctypes.byref(self.d), shim._void_array([t.data_ptr() for t in self.inputs]),
          shim._void_array([self.out.data_ptr()]), self.ws.data_ptr(), nbytes, self._stream(),
          ctypes.byref(self.session))
    if rc != 0:
      raise RuntimeError(f"b200fdtd_session_create failed ({rc}): {shim._last_error()}")
    info = (ctypes.c_int64 * 8)()
    L.b200fdtd_session_layout(self.session, info)
    e_off, h_off, comp_b, plane_b, zp, el, X, Y = (int(v) for v in info)
    dt = torch.float32 if el == 4 else torch.float16

    def view(off):
      flat = self.ws[off:off + 3 * comp_b].view(dt)
      return flat.view(3, X, Y, zp)
    self.E, self.H = view(e_off), view(h_off)                # zero-copy views (3, X, Y, Zp)

  def _stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def step_h(self):
    rc = self.L.b200fdtd_session_step_h(self.session, self._stream())
    if rc:
      raise RuntimeError(self.shim._last_error())

  def step_e(self, n):
    rc = self.L.b200fdtd_session_step_e(self.session, int(n), self._stream())
    if rc:
      raise RuntimeError(self.shim._last_error())

  def snapshots(self):
    return self.out

  def close(self):
    if self.session:
      torch.cuda.synchronize(self.device)
      self.L.b200fdtd_session_destroy(self.session)
      self.session = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


def _exchange(send, recv, to_rank, from_rank, group, world):
  """send -> rank `to_rank`'s recv buffer; fills `recv` from `from_rank`."""
  if world == 1:
    recv.copy_(send)
    return
  ops = [dist.P2POp(dist.isend, send, to_rank, group), dist.P2POp(dist.irecv, recv, from_rank, group)]
  for w in dist.batch_isend_irecv(ops):
    w.wait()


class DecomposedRun:
  """One x-decomposed engine call, split into set-up (``__init__``: slab inputs, session,
  coefficient preparation) and the time loop (``run``), so benchmarks can time the loop alone."""

  def __init__(self, kw, group=None, make_slab=None, device=None):
    self.kw, self.group = kw, group
    self.world, self.rank = 1, 0
    if dist.is_available() and dist.is_initialized():
      self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
    loc, self.nloc, self.crop = local_problem(kw, self.rank, self.world)
    if make_slab is None:
      if not torch.cuda.is_available():
        raise RuntimeError("fdtdz_decomposed needs a CUDA device (no CPU fallback)")
      dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
      self.slab = CudaSlab(loc, dev)
    else:
      self.slab = make_slab(loc)
    self.tt = _np(kw["source_waveform"]).shape[0]

  def _peer(self, r):
    r %= self.world
    if self.group is not None and self.world > 1:
      return dist.get_global_rank(self.group, r)
    return r

  def run(self):
    slab, nloc, group, world = self.slab, self.nloc, self.group, self.world
    right, left = self._peer(self.rank + 1), self._peer(self.rank - 1)
    E, H = slab.E, slab.H                                    # (3, nloc+2, Y, Zp)
    face = torch.empty_like(H[:, 0].contiguous())            # packed (3, Y, Zp) receive buffer
    for n in range(self.tt):
      slab.step_h()
      # my last owned H plane -> right neighbour's low ghost; my low ghost <- left neighbour
      _exchange(H[:, nloc].contiguous(), face, right, left, group, world)
      H[:, 0].copy_(face)
      slab.step_e(n)
      # my first owned E plane -> left neighbour's high ghost; my high ghost <- right neighbour
      _exchange(E[:, 1].contiguous(), face, left, right, group, world)
      E[:, nloc + 1].copy_(face)

  def local_snapshots(self):
    """(x_lo, x_hi, snapshots[:, :, x_lo:x_hi]) of this rank, in output coordinates."""
    snaps = self.slab.snapshots()                            # (n_out, 3, nloc+2, yy, zz)
    if not isinstance(snaps, torch.Tensor):
      snaps = torch.from_numpy(np.ascontiguousarray(snaps))
    if self.crop is None:
      return 0, 0, snaps[:, :, :0]
    return self.crop[2], self.crop[3], snaps[:, :, self.crop[0]:self.crop[1]]

  def gathered_snapshots(self):
    lo, hi, snaps = self.local_snapshots()
    shape = tuple(self.kw["epsilon"].shape)
    full = torch.zeros((snaps.shape[0], 3) + shape[1:], dtype=torch.float32, device=snaps.device)
    if hi > lo:
      full[:, :, lo:hi] = snaps
    if self.world > 1:
      dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)   # disjoint slabs: sum == concat
    return full

  def close(self):
    if hasattr(self.slab, "close"):
      self.slab.close()


def fdtdz_decomposed(epsilon, dt, source_field, source_waveform, source_position,
                     absorption_mask, pml_kappa, pml_sigma, pml_alpha, pml_widths,
                     output_steps, use_reduced_precision, launch_params=None, offset=(0, 0, 0),
                     *, group=None, make_slab=None, device=None, gather=True):
  """The engine call, x-decomposed over the ranks of ``group`` (default: the world group).

  Every rank passes the SAME global arguments.  Returns the global ``(n_out, 3, xx, yy, zz)``
  snapshots on every rank (``gather=True``) or this rank's ``(x_lo, x_hi, local snapshots)``.
  ``make_slab(local_kwargs)`` builds the slab engine (default: ``CudaSlab`` on ``device`` /
  the current CUDA device); the CPU tests pass an oracle-backed one.
  """
  kw = dict(epsilon=epsilon, dt=dt, source_field=source_field, source_waveform=source_waveform,
            source_position=source_position, absorption_mask=absorption_mask,
            pml_kappa=pml_kappa, pml_sigma=pml_sigma, pml_alpha=pml_alpha,
            pml_widths=pml_widths, output_steps=output_steps,
            use_reduced_precision=use_reduced_precision, launch_params=launch_params,
            offset=offset)
  run = DecomposedRun(kw, group=group, make_slab=make_slab, device=device)
  run.run()
  out = run.gathered_snapshots() if gather else run.local_snapshots()
  if gather:
    run.close()
  return out
