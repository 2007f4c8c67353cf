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
    # Ownership is decided on GLOBAL plane indices (a local modulo would map plane X-1 of a slab
    # that covers the periodic wrap onto the low ghost instead of its owned copy).
    g0, g1 = p % X, (p - 1) % X
    own0, own1 = x0 <= g0 < x1, x0 <= g1 < x1
    wf = wf.copy()
    if not own0:
      wf[:, 0] = 0
    if not own1:
      wf[:, 1] = 0
    if own0 and own1:
      lp = g0 - x0 + 1
      if g1 - x0 + 1 != lp - 1 and wf[:, 1].any():
        # the engine injects channel 1 on the plane before channel 0's; here the two owned planes
        # are not adjacent in local coordinates (the slab covers the periodic wrap)
        raise NotImplementedError("x source at plane 0 of a slab that also owns plane X-1")
    elif own0:
      lp = g0 - x0 + 1                                       # channel 1 (zeroed) lands on lp-1
    elif own1:
      lp = g1 - x0 + 2                                       # channel 0 (zeroed) lands on lp
    else:
      lp = 1
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
    self.inputs = [loc[k].to(self.device, torch.float32).contiguous() if isinstance(loc[k], torch.Tensor)
                   else torch.from_numpy(np.ascontiguousarray(_np(loc[k]), np.float32)).to(self.device)
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


# ===================================================================================================
# y-slab decomposition with ghost zones: G time steps per halo exchange, systolic kernel per slab
# ===================================================================================================
#
# The x-slab scheme above exchanges a face every half-step, which ties it to the per-step kernels.
# Cutting along y instead leaves the x sweep of the persistent systolic kernel intact: every rank
# owns Y/N columns plus G ghost columns per side, copies its neighbours' edge columns (E, H and
# the CPML psi arrays) into the ghosts, and then advances G whole steps in ONE launch
# (``b200fdtd_session_advance``) on its local, periodically wrapped domain.  Wrong data enters
# only through the wrap at the outer ghost edge and travels one column per step, so after G steps
# it has consumed exactly the ghost zone: owned cells see the inputs of the single-GPU run and the
# result is bit-identical.  Cost: (Yo + 2G) / Yo redundant work and one packed exchange per G steps.


def local_problem_y(kw, rank, world, ghost):
  """Engine kwargs of this rank's y-slab (ghost columns included) + snapshot crop."""
  eps = _np(kw["epsilon"]).astype(np.float32)
  mask = _np(kw["absorption_mask"]).astype(np.float32)
  sf = _np(kw["source_field"]).astype(np.float32)
  Y = mask.shape[2]
  ox, oy, oz = (int(o) for o in kw["offset"])
  yy = eps.shape[2]
  y0, y1 = slab_bounds(Y, world, rank)
  nloc = y1 - y0
  G = int(ghost)
  if nloc < max(G, 1):
    raise ValueError(f"slab of {nloc} columns is narrower than the ghost zone ({G})")
  cols = np.arange(y0 - G, y1 + G) % Y                       # global column of each local column
  eps_y = np.clip(cols - oy, 0, yy - 1)                      # edge replication in y
  loc = dict(kw)
  loc["epsilon"] = np.ascontiguousarray(eps[:, :, eps_y])
  loc["absorption_mask"] = np.ascontiguousarray(mask[:, :, cols])
  loc["offset"] = (ox, 0, oz)
  axis = 2 if sf.ndim == 5 else (0 if sf.shape[1] == 1 else 1)
  wf = _np(kw["source_waveform"]).astype(np.float32)
  if axis == 0:
    loc["source_field"] = np.ascontiguousarray(sf[:, :, cols])
  elif axis == 2:
    loc["source_field"] = np.ascontiguousarray(sf[:, :, :, cols])
  else:
    # y-plane source: channel 0 acts on global column p, channel 1 on p-1.  Both must be
    # applied wherever they fall inside the local range (ghost cells are recomputed, too); at
    # the outermost ghost column the wrapped partner lands on the other edge column, which is
    # already invalid after the first step.
    p = int(kw["source_position"])
    hit = np.nonzero(cols == p % Y)[0]
    hit = [int(h) for h in hit if h >= 1]
    if len(hit) > 1:
      # (only when a slab plus its ghosts is wider than the domain, e.g. a single rank)
      raise NotImplementedError("y-plane source visible twice in one slab (owned column and "
                                "ghost image): use fewer ghost columns or more ranks")
    if hit:
      loc["source_position"] = hit[0]
    else:
      loc["source_position"] = 1
      wf = np.zeros_like(wf)
    loc["source_field"] = sf
  loc["source_waveform"] = wf
  g0, g1 = max(oy, y0), min(oy + yy, y1)
  crop = None
  if g1 > g0:
    crop = (g0 - y0 + G, g1 - y0 + G, g0 - oy, g1 - oy)      # (local lo, local hi, out lo, out hi)
  return loc, nloc, crop


def choose_ghost(kw, world, min_ghost=None):
  """Ghost width for the y-slab scheme: the smallest multiple of the systolic kernel's stage
  count >= ``min_ghost`` (a launch of G steps keeps all S pipeline stages busy only if S divides
  G).  ``min_ghost`` defaults to 16 for slabs of >= 384 columns (halves the exchange count for
  3 % more redundant work; measured +3 % on 512-column slabs) and 8 below.  A pure function of
  the global shapes (widest slab), so every rank picks the same value."""
  from . import fdtdz_jax as shim
  mask = kw["absorption_mask"]
  eps = kw["epsilon"]
  sf = kw["source_field"]
  X, Y = int(mask.shape[1]), int(mask.shape[2])
  nloc = -(-Y // world)
  if min_ghost is None:
    min_ghost = 16 if nloc >= 384 else 8
  axis = 2 if len(sf.shape) == 5 else (0 if sf.shape[1] == 1 else 1)
  Z = int(np.asarray(kw["pml_kappa"]).shape[0])
  zero = np.float32(0)
  G = int(min_ghost)
  for _ in range(4):
    Yl = nloc + 2 * G
    loc = dict(kw)
    loc["epsilon"] = np.broadcast_to(zero, (3, int(eps.shape[1]), Yl, int(eps.shape[3])))
    loc["absorption_mask"] = np.broadcast_to(zero, (3, X, Yl))
    loc["source_field"] = np.broadcast_to(zero, {0: (2, 1, Yl, Z), 1: (2, X, 1, Z),
                                                 2: (2, 2, X, Yl, 1)}[axis])
    loc["offset"] = (int(kw["offset"][0]), 0, int(kw["offset"][2]))
    loc["source_position"] = min(int(kw["source_position"]), (Yl if axis == 1 else 10**9) - 1)
    info = shim.plan_info(**loc)
    S = info["stages"] if info["kernel"] == "systolic_lean" else 1
    G2 = S * max(1, -(-int(min_ghost) // S))
    if G2 == G:
      break
    G = G2
  return min(G, Y // world)


class CudaSlabY:
  """y-slab engine over the C ABI's stepping session: whole-step ``advance`` (one persistent
  systolic launch per call when the geometry allows it, per-step kernels otherwise)."""

  def __init__(self, loc, device, kernel="auto"):
    from . import fdtdz_jax as shim
    self.shim = shim
    L = shim.lib()
    loc = dict(loc)
    lp = dict(loc.get("launch_params") or {})
    lp.setdefault("kernel", kernel)
    loc["launch_params"] = lp
    self.d = shim.make_desc(**{k: loc[k] for k in (
        "epsilon", "dt", "source_field", "source_waveform", "source_position", "absorption_mask",
        "pml_kappa", "pml_sigma", "pml_alpha", "pml_widths", "output_steps",
        "use_reduced_precision", "launch_params", "offset")})
    self.device = torch.device(device)
    names = ("epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa",
             "pml_sigma", "pml_alpha")
    self.inputs = [loc[k].to(self.device, torch.float32).contiguous() if isinstance(loc[k], torch.Tensor)
                   else torch.from_numpy(np.ascontiguousarray(_np(loc[k]), np.float32)).to(self.device)
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
    L.b200fdtd_session_advance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p]
    L.b200fdtd_session_layout2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
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
    info = (ctypes.c_int64 * 16)()
    L.b200fdtd_session_layout2(self.session, info)
    (e_off, h_off, comp_b, plane_b, zp, el, X, Y, e2_off, psi_off, psi_b, psi2_off, psi_k,
     pingpong, kernel_id, stages) = (int(v) for v in info)
    self.pingpong, self.stages = bool(pingpong), stages
    self.kernel = {v: k for k, v in shim._KERNELS.items()}.get(kernel_id, str(kernel_id))
    dt = torch.float32 if el == 4 else torch.float16

    def fview(off):
      return self.ws[off:off + 3 * comp_b].view(dt).view(3, X, Y, zp)

    def pview(off):
      return self.ws[off:off + 2 * psi_b].view(torch.float32).view(2, X, Y, psi_k)
    psiE = pview(psi_off + 2 * psi_b)
    self._sets = [[fview(e_off), fview(h_off), pview(psi_off), psiE]]
    if self.pingpong:
      self._sets.append([fview(e2_off), fview(e2_off + (h_off - e_off)), pview(psi2_off), psiE])

  def _stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def state(self, n):
    """[E, H, psiH, psiE] views (C, X, Y, .) holding the state after ``n`` steps."""
    return self._sets[n & 1] if self.pingpong else self._sets[0]

  def advance(self, n0, nsteps):
    rc = self.L.b200fdtd_session_advance(self.session, int(n0), int(nsteps), self._stream())
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


def _pack(state, lo, hi):
  """Columns [lo, hi) of every state array, as one contiguous byte buffer."""
  return torch.cat([t[:, :, lo:hi].reshape(-1).view(torch.uint8) for t in state])


def _unpack(buf, state, lo, hi):
  off = 0
  for t in state:
    dst = t[:, :, lo:hi]
    n = dst.numel() * dst.element_size()
    dst.copy_(buf[off:off + n].view(dst.dtype).view(dst.shape))
    off += n


class YSlabRun:
  """One y-decomposed engine call: set-up in ``__init__``, the time loop in ``run``."""

  def __init__(self, kw, ghost=8, group=None, make_slab=None, device=None, kernel="auto",
               local=None, solo=False):
    """``local`` = a pre-built ``(local kwargs, nloc, crop)`` triple (what ``local_problem_y``
    returns) for domains whose global arrays are too large to materialise on every rank.
    ``solo``: wrap the slab onto itself on this rank alone, whatever process group exists."""
    self.kw, self.group = kw, group
    self.world, self.rank = 1, 0
    if not solo and dist.is_available() and dist.is_initialized():
      self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
    self.G = choose_ghost(kw, self.world) if ghost is None else int(ghost)
    loc, self.nloc, self.crop = (local if local is not None else
                                 local_problem_y(kw, self.rank, self.world, self.G))
    if make_slab is None:
      if not torch.cuda.is_available():
        raise RuntimeError("fdtdz_decomposed_y needs a CUDA device (no CPU fallback)")
      dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
      self.slab = CudaSlabY(loc, dev, kernel)
    else:
      self.slab = make_slab(loc)
    self.tt = _np(kw["source_waveform"]).shape[0]

  def _peer(self, r):
    r %= self.world
    if self.group is not None and self.world > 1:
      return dist.get_global_rank(self.group, r)
    return r

  def run(self):
    slab, G, group, world = self.slab, self.G, self.group, self.world
    right, left = self._peer(self.rank + 1), self._peer(self.rank - 1)
    Yl = self.nloc + 2 * G
    timing = getattr(self, "time_exchange", False) and torch.cuda.is_available()
    marks = []
    for n0 in range(0, self.tt, max(G, 1)):
      st = slab.state(n0)
      if n0 > 0:                                             # (the initial state is all zero)
        if timing:
          a = torch.cuda.Event(enable_timing=True); a.record()
        send_lo = _pack(st, G, 2 * G)                        # my low owned edge
        send_hi = _pack(st, Yl - 2 * G, Yl - G)              # my high owned edge
        recv_hi, recv_lo = torch.empty_like(send_lo), torch.empty_like(send_hi)
        _exchange(send_lo, recv_hi, left, right, group, world)   # -> left's high ghost
        _exchange(send_hi, recv_lo, right, left, group, world)   # -> right's low ghost
        _unpack(recv_hi, st, Yl - G, Yl)
        _unpack(recv_lo, st, 0, G)
        if timing:
          b = torch.cuda.Event(enable_timing=True); b.record()
          marks.append((a, b))
      slab.advance(n0, min(G, self.tt - n0))
    if timing:
      torch.cuda.synchronize()
      self.exchange_ms = sum(a.elapsed_time(b) for a, b in marks)   # pack + send/recv + unpack

  def local_snapshots(self):
    snaps = self.slab.snapshots()                            # (n_out, 3, xx, nloc+2G, zz)
    if not isinstance(snaps, torch.Tensor):
      snaps = torch.from_numpy(np.ascontiguousarray(snaps))
    if self.crop is None:
      return 0, 0, snaps[:, :, :, :0]
    return self.crop[2], self.crop[3], snaps[:, :, :, self.crop[0]:self.crop[1]]

  def gathered_snapshots(self):
    lo, hi, snaps = self.local_snapshots()
    shape = tuple(self.kw["epsilon"].shape)
    full = torch.zeros((snaps.shape[0], 3) + shape[1:], dtype=torch.float32, device=snaps.device)
    if hi > lo:
      full[:, :, :, lo:hi] = snaps
    if self.world > 1:
      dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
    return full

  def close(self):
    if hasattr(self.slab, "close"):
      self.slab.close()


def fdtdz_decomposed_y(epsilon, dt, source_field, source_waveform, source_position,
                       absorption_mask, pml_kappa, pml_sigma, pml_alpha, pml_widths,
                       output_steps, use_reduced_precision, launch_params=None, offset=(0, 0, 0),
                       *, ghost=8, group=None, make_slab=None, device=None, gather=True):
  """The engine call, y-decomposed with ``ghost`` ghost columns per side over the ranks of
  ``group``: ``ghost`` time steps per halo exchange (``None``: the smallest multiple of the
  kernel's pipeline depth >= 8).  Same contract as ``fdtdz_decomposed``."""
  kw = dict(epsilon=epsilon, dt=dt, source_field=source_field, source_waveform=source_waveform,
            source_position=source_position, absorption_mask=absorption_mask,
            pml_kappa=pml_kappa, pml_sigma=pml_sigma, pml_alpha=pml_alpha,
            pml_widths=pml_widths, output_steps=output_steps,
            use_reduced_precision=use_reduced_precision, launch_params=launch_params,
            offset=offset)
  run = YSlabRun(kw, ghost=ghost, group=group, make_slab=make_slab, device=device)
  run.run()
  out = run.gathered_snapshots() if gather else run.local_snapshots()
  if gather:
    run.close()
  return out


# ===================================================================================================
# y-slab decomposition with IN-KERNEL halo exchange: peer-mapped stores over NVLink, one launch
# ===================================================================================================
#
# The ghost-zone scheme above pays (Yo + 2G) / Yo redundant work, one host-driven NCCL round per G
# steps and a pipeline fill/drain per launch.  Here every rank keeps ONE ghost column per side and
# the persistent launch itself does the exchange (include/b200fdtd.h, "y-slab sessions"): a courier
# CTA copies each newly finished plane of the slab's edge columns into the neighbour's ghost column
# through peer-mapped memory, then the edge tile's progress counter into the neighbour's mirror
# slot, and the neighbour's edge tiles wait on it like on any local tile.  The whole run is one launch per
# GPU; transfers overlap the interior update tile by tile; no redundant cells.  Bit-identical to the
# single-GPU run by construction (every owned cell sees the same operands in the same order).


class CudaSlabP2P:
  """y-slab engine over ``b200fdtd_session_create_slab`` + peer-mapped workspaces."""

  def __init__(self, loc, device, group=None, solo=False):
    """``solo``: the slab's neighbours are the slab itself on this rank alone (no collectives),
    whatever process group exists -- the 1-GPU reference of a weak-scaling measurement."""
    from . import fdtdz_jax as shim
    self.shim, self.group = shim, group
    L = self.L = shim.lib()
    self.world, self.rank = 1, 0
    if not solo and dist.is_available() and dist.is_initialized():
      self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
    loc = dict(loc)
    lp = dict(loc.get("launch_params") or {})
    lp["kernel"] = "systolic_lean"
    loc["launch_params"] = lp
    self.d = shim.make_desc(**{k: loc[k] for k in (
        "epsilon", "dt", "source_field", "source_waveform", "source_position", "absorption_mask",
        "pml_kappa", "pml_sigma", "pml_alpha", "pml_widths", "output_steps",
        "use_reduced_precision", "launch_params", "offset")})
    self.device = torch.device(device)
    names = ("epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa",
             "pml_sigma", "pml_alpha")
    self.inputs = [loc[k].to(self.device, torch.float32).contiguous() if isinstance(loc[k], torch.Tensor)
                   else torch.from_numpy(np.ascontiguousarray(_np(loc[k]), np.float32)).to(self.device)
                   for k in names]
    Y = self.d.Y
    L.b200fdtd_session_workspace_bytes_slab.restype = ctypes.c_size_t
    L.b200fdtd_session_workspace_bytes_slab.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    nbytes = L.b200fdtd_session_workspace_bytes_slab(ctypes.byref(self.d), 1, Y - 1)
    if nbytes == 0:
      raise RuntimeError(shim._last_error())
    self.nbytes = nbytes
    L.b200fdtd_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    L.b200fdtd_peer_free.argtypes = [ctypes.c_void_p]
    L.b200fdtd_peer_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_peer_open.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_peer_close.argtypes = [ctypes.c_void_p]
    self.ws = ctypes.c_void_p()
    self._opened = []
    with torch.cuda.device(self.device):
      self._check(L.b200fdtd_peer_alloc(nbytes, ctypes.byref(self.ws)))
      nout = L.b200fdtd_num_outputs(ctypes.byref(self.d))
      self.out = torch.zeros((nout, 3, self.d.xx, self.d.yy, self.d.zz), dtype=torch.float32,
                             device=self.device)
      self.session = ctypes.c_void_p()
      L.b200fdtd_session_create_slab.argtypes = [
          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
          ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
      self._check(L.b200fdtd_session_create_slab(
          ctypes.byref(self.d), shim._void_array([t.data_ptr() for t in self.inputs]),
          shim._void_array([self.out.data_ptr()]), self.ws, nbytes, self._stream(), 1, Y - 1,
          ctypes.byref(self.session)))
      lo, hi = self._map_neighbours()
      L.b200fdtd_session_set_peers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
      self._check(L.b200fdtd_session_set_peers(self.session, lo, hi))
    L.b200fdtd_session_advance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p]
    L.b200fdtd_session_slab_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_session_destroy.argtypes = [ctypes.c_void_p]
    L.b200fdtd_session_destroy.restype = None
    info = (ctypes.c_int64 * 16)()
    L.b200fdtd_session_layout2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.b200fdtd_session_layout2(self.session, info)
    self.stages = int(info[15])
    self.kernel = "systolic_lean"

  def _check(self, rc):
    if rc:
      raise RuntimeError(self.shim._last_error())

  def _stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def _map_neighbours(self):
    """Device addresses of the low / high neighbour's workspace in THIS process (CUDA IPC)."""
    if self.world == 1:
      return self.ws, self.ws                                # the slab wraps onto itself
    # Every step below is collective-safe: a failure on one rank is agreed on by all of them
    # before anybody raises, so that the caller can fall back to the NCCL path in step.
    handle = (ctypes.c_ubyte * 64)()
    ok = self.L.b200fdtd_peer_export(self.ws, handle) == 0
    err = "" if ok else self.shim._last_error()
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
    allh = [torch.empty_like(mine) for _ in range(self.world)]
    dist.all_gather(allh, mine, group=self.group)
    shapes = torch.tensor([int(ok), self.nbytes, self.d.X, self.d.Y, self.d.Z], dtype=torch.int64,
                          device=self.device)
    alls = [torch.empty_like(shapes) for _ in range(self.world)]
    dist.all_gather(alls, shapes, group=self.group)
    if any(int(a[0]) == 0 for a in alls):
      raise RuntimeError(f"CUDA IPC export failed on some rank: {err}")
    if any(not torch.equal(a, shapes) for a in alls):
      raise ValueError("in-kernel halo exchange needs the same local shape on every rank: "
                       f"{[a.tolist() for a in alls]}")
    ptrs = {self.rank: self.ws}
    for r in {(self.rank - 1) % self.world, (self.rank + 1) % self.world} - {self.rank}:
      hb = (ctypes.c_ubyte * 64)(*allh[r].cpu().tolist())
      p = ctypes.c_void_p()
      if self.L.b200fdtd_peer_open(hb, ctypes.byref(p)) == 0:
        self._opened.append(p)
        ptrs[r] = p
      else:
        ok, err = False, self.shim._last_error()
    flag = torch.tensor([int(ok)], device=self.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
    if int(flag.item()) == 0:
      raise RuntimeError(f"CUDA IPC mapping of a neighbour's workspace failed on some rank: {err}")
    return ptrs[(self.rank - 1) % self.world], ptrs[(self.rank + 1) % self.world]

  def _barrier(self):
    torch.cuda.synchronize(self.device)
    if self.world > 1:
      dist.barrier(group=self.group)

  def advance(self, n0, nsteps):
    self._barrier()                      # every rank has finished its previous launch
    with torch.cuda.device(self.device):
      self._check(self.L.b200fdtd_session_slab_reset(self.session, self._stream()))
    self._barrier()                      # every rank's counters and mirror slots are clear
    with torch.cuda.device(self.device):
      self._check(self.L.b200fdtd_session_advance(self.session, int(n0), int(nsteps), self._stream()))

  def snapshots(self):
    return self.out

  def close(self):
    if getattr(self, "session", None):
      self._barrier()                    # nobody is still storing into this workspace
      self.L.b200fdtd_session_destroy(self.session)
      self.session = ctypes.c_void_p()
      for p in self._opened:
        self.L.b200fdtd_peer_close(p)
      self._opened = []
      if self.world > 1:
        dist.barrier(group=self.group)   # every mapping is closed before the memory goes away
      self.L.b200fdtd_peer_free(self.ws)
      self.ws = ctypes.c_void_p()

  def __del__(self):
    try:
      if getattr(self, "session", None) and self.world == 1:
        self.close()
    except Exception:
      pass


class P2PSlabRun:
  """One y-decomposed engine call with in-kernel halo exchange: set-up in ``__init__``, the time
  loop (ONE persistent launch per GPU) in ``run``.  Same interface as ``YSlabRun``."""

  def __init__(self, kw, group=None, device=None, local=None, solo=False):
    self.kw, self.group = kw, group
    self.world, self.rank = 1, 0
    if not solo and dist.is_available() and dist.is_initialized():
      self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
    self.G = 1
    loc, self.nloc, self.crop = (local if local is not None else
                                 local_problem_y(kw, self.rank, self.world, 1))
    if not torch.cuda.is_available():
      raise RuntimeError("P2PSlabRun needs a CUDA device (no CPU fallback)")
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    self.slab = CudaSlabP2P(loc, dev, group, solo=solo)
    self.tt = _np(loc["source_waveform"]).shape[0]

  def run(self):
    self.slab.advance(0, self.tt)

  local_snapshots = YSlabRun.local_snapshots
  gathered_snapshots = YSlabRun.gathered_snapshots

  def close(self):
    self.slab.close()


def fdtdz_decomposed_p2p(epsilon, dt, source_field, source_waveform, source_position,
                         absorption_mask, pml_kappa, pml_sigma, pml_alpha, pml_widths,
                         output_steps, use_reduced_precision, launch_params=None, offset=(0, 0, 0),
                         *, group=None, device=None, gather=True):
  """The engine call, y-decomposed with in-kernel halo exchange over peer-mapped memory.  Same
  contract as ``fdtdz_decomposed_y``; needs fp32 storage, 125 <= Z <= 128 and Y divisible by the
  number of ranks."""
  kw = dict(epsilon=epsilon, dt=dt, source_field=source_field, source_waveform=source_waveform,
            source_position=source_position, absorption_mask=absorption_mask,
            pml_kappa=pml_kappa, pml_sigma=pml_sigma, pml_alpha=pml_alpha,
            pml_widths=pml_widths, output_steps=output_steps,
            use_reduced_precision=use_reduced_precision, launch_params=launch_params,
            offset=offset)
  run = P2PSlabRun(kw, group=group, device=device)
  try:
    run.run()
    out = run.gathered_snapshots() if gather else run.local_snapshots()
  except BaseException:
    if run.world == 1:                   # (with neighbours, close() is collective: a rank that
      run.close()                        # failed alone must not wait for the others in it)
    raise
  run.close()
  return out


def decomposed_engine(kind="p2p", **opts):
  """An ``engine=`` for ``pjz_b200.field`` / ``scatter`` that solves EVERY engine call on ALL ranks
  of the process group, the domain cut into slabs (BASELINE.json config 5: domains too large or
  too slow for one GPU), instead of dealing whole ports to the ranks.

  ``kind``: "p2p" (y-slabs, halo exchange inside the persistent kernel), "y" (y-slabs with ghost
  zones over NCCL / gloo) or "x" (x-slabs, exchange per half-step).  ``opts`` go to the driver
  (``group``, ``device``, ``ghost``, ``make_slab``).  The callable is marked ``collective``:
  ``scatter`` then runs every port on every rank and skips its own broadcast of the fields.
  """
  fn = {"p2p": fdtdz_decomposed_p2p, "y": fdtdz_decomposed_y, "x": fdtdz_decomposed}[kind]

  def engine(**kw):
    if kw.pop("output_projection", None) is not None:
      raise NotImplementedError("the decomposed engines return snapshots; use fuse_projection=False")
    return fn(**kw, **opts)

  engine.collective = True
  engine.__name__ = f"fdtdz_decomposed_{kind}"
  return engine
