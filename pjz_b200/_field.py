"""Host-side mirror of pjz's time-harmonic glue around the FDTD engine call.

Restates /root/reference/src/pjz/_field.py (``SimParams`` :12-53, input builders :56-167,
``field`` :171-279, ``scatter`` :282-442) with NumPy for the small host tables and torch for
the volume-sized tensors, so the one hot call -- ``fdtdz_jax.fdtdz(**14 kwargs)``,
/root/reference/src/pjz/_field.py:254-269 -- can be driven exactly the way pjz drives it.
JAX is not part of this image, so ``jax.jit``/``custom_vjp`` become plain functions and a
``torch.autograd.Function``.

The engine defaults to the CUDA implementation (``pjz_b200.fdtdz_jax.fdtdz``); there is no
CPU fallback.  ``engine=`` exists so tests can run the identical glue over the oracle.
"""

from __future__ import annotations

from typing import Any, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch


class SimParams(NamedTuple):
  """Mirror of ``pjz.SimParams`` (/root/reference/src/pjz/_field.py:12-53), same fields,
  order and defaults.  ``domain_zz`` is the one extension: pjz fixes the engine height at
  128 (reduced precision) / 64 (fp32) cells *including the PML bookkeeping*
  (``_zz``, :56-58) because fdtd-z maps z onto one warp; this engine accepts any height, so
  ``domain_zz`` (total z cells of the engine domain, PML included) overrides that rule when
  set.  ``None`` reproduces pjz exactly.
  """
  omega_range: Tuple[float, float]
  tt: int
  dt: float = 0.5
  source_ramp: float = 4.0
  source_delay: float = 4.0
  absorption_padding: int = 50
  absorption_coeff: float = 1e-4
  pml_widths: Tuple[int, int] = (16, 16)
  pml_alpha_coeff: float = 0.0
  pml_sigma_lnr: float = 0.5
  pml_sigma_m: float = 1.3
  use_z_as_batch: bool = False
  use_reduced_precision: bool = True
  launch_params: Any = None
  domain_zz: Optional[int] = None


def _zz(pml_widths, use_reduced_precision, domain_zz=None):
  """Height of the engine domain (/root/reference/src/pjz/_field.py:56-58)."""
  if domain_zz is not None:
    return int(domain_zz)
  return (128 if use_reduced_precision else 64) - sum(pml_widths)


def _pad_zz(epsilon_zz, pml_widths, use_reduced_precision, domain_zz=None):
  """z padding that centres epsilon in the engine domain (:61-66)."""
  zz = _zz(pml_widths, use_reduced_precision, domain_zz)
  bot = (zz - epsilon_zz) // 2
  top = zz - epsilon_zz - bot
  if bot < 0 or top < 0:
    raise ValueError(f"epsilon height {epsilon_zz} exceeds the engine domain height {zz}")
  return (bot, top)


def _absorption_profiles(numcells, width, smoothness):
  """1-D quadratic absorber profiles at offsets 0 and 1/2 (:69-76)."""
  center = (numcells - 1) / 2
  offset = np.array([[0.0], [0.5]])
  pos = np.arange(numcells) + offset
  pos = np.abs(pos - center) - center + width
  pos = np.clip(pos, 0, None)
  return smoothness * np.power(pos, 2)


def _cross_profiles(x, y):
  return np.maximum(*np.meshgrid(x, y, indexing="ij"))[None, ...]


def _absorption_mask(xx, yy, width, smoothness):
  """``(3, xx, yy)`` absorber conductivity (:84-90; KAT tests/test_boundaries.py:8-24)."""
  x = _absorption_profiles(xx, width, smoothness)
  y = _absorption_profiles(yy, width, smoothness)
  return np.concatenate([_cross_profiles(x[0], y[1]),
                         _cross_profiles(x[1], y[0]),
                         _cross_profiles(x[1], y[1])]).astype(np.float32)


def _safe_div(x, y):
  return np.zeros_like(x) if y == 0 else x / y


def _pml_sigma(pml_widths, zz, ln_R, m):
  """``(zz, 2)`` PML conductivity (:98-106; KAT tests/test_boundaries.py:27-41)."""
  offset = np.array([[0.0], [0.5]])
  z = np.arange(zz) + offset
  z = np.stack([_safe_div(pml_widths[0] - z, pml_widths[0]),
                _safe_div(z + 0.5 - zz + pml_widths[1], pml_widths[1])], axis=-1)
  z = np.max(np.clip(z, 0, None), axis=-1)
  return ((m + 1) * ln_R * z**m).T.astype(np.float32)


def _ramped_sin(omega, width, delay, dt, tt):
  """Mean over omega of a tanh-ramped complex exponential (:109-113)."""
  omega = np.asarray(omega, np.float64).reshape(-1)
  t = omega[:, None] * dt * np.arange(tt)
  waveforms = ((1 + np.tanh(t / width - delay)) / 2) * np.exp(1j * t)
  return np.mean(waveforms, axis=0)


def _sampling_interval(omega_min, omega_max, omega_n, dt):
  """Snapshot spacing that keeps ``omega_n`` components observable (:116-139)."""
  period = 4 * np.pi / (omega_max + omega_min) / dt
  if omega_n == 1:
    return int(round(period / 4))
  cutoff = np.pi * (omega_n - 1) / omega_n / (omega_max - omega_min) / dt
  m = np.floor((cutoff - period / 4) / (period / 2))
  return int(round(period / 4 + m * period / 2))


def _output_phases(omega, output_steps, dt):
  """``(2ww, n_out)`` rows cos(w t_n) then -sin(w t_n) (:142-150)."""
  omega = np.asarray(omega, np.float64).reshape(-1)
  steps = np.arange(*output_steps)
  theta = omega[:, None] * dt * steps
  return np.concatenate([np.cos(theta), -np.sin(theta)], axis=0)


def _prop_axis(mode):
  """Propagation axis from the singleton spatial dimension (:282-290)."""
  shp = tuple(mode.shape[-3:])
  if shp.count(1) == 1:
    return "xyz"[shp.index(1)]
  raise ValueError(
      f"``mode.shape[-3:]`` must contain exactly one value of ``1``, "
      f"instead got ``mode.shape == {tuple(mode.shape)}``.")


def _transverse_slice(arr, pos, axis):
  """Two transverse components of ``arr`` (..., 3, X, Y, Z) on plane ``pos`` (:153-163)."""
  a = "xyz".find(axis)
  comps = [i for i in range(3) if i != a]
  arr = arr[..., comps, :, :, :]
  return arr.narrow(arr.ndim - 3 + a, int(pos), 1)


def _source(mode, pos, epsilon):
  return mode / _transverse_slice(epsilon, pos, _prop_axis(mode))


def _default_engine():
  from . import fdtdz_jax
  return fdtdz_jax.fdtdz


def _as_tensor(a, device=None, dtype=None):
  if isinstance(a, torch.Tensor):
    t = a
  else:
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
  if dtype is not None and t.dtype != dtype:
    t = t.to(dtype)
  if device is not None and t.device != torch.device(device):
    t = t.to(device)
  return t


def engine_inputs(epsilon, source, omega, source_pos, sim_params):
  """Everything ``field`` hands to the engine, as a kwargs dict (mirrors :194-269).

  Small tables are NumPy float32; ``epsilon`` and ``source_field`` are torch tensors on
  epsilon's device.  Also returns ``output_steps``-derived data needed afterwards.
  """
  (omega_range, tt, dt, source_ramp, source_delay, absorption_padding, absorption_coeff,
   pml_widths, pml_alpha_coeff, pml_sigma_lnr, pml_sigma_m, use_z_as_batch,
   use_reduced_precision, launch_params, domain_zz) = sim_params
  epsilon = _as_tensor(epsilon, dtype=torch.float32)
  dev = epsilon.device
  omega_np = (omega.detach().cpu().numpy() if isinstance(omega, torch.Tensor)
              else np.asarray(omega)).astype(np.float64).reshape(-1)
  source = _as_tensor(source, dev)
  if epsilon.ndim != 4 or epsilon.shape[0] != 3:
    raise ValueError(f"epsilon must be (3, xx, yy, zz), got {tuple(epsilon.shape)}")
  if source.ndim != 4 or source.shape[0] != 2:
    raise ValueError(f"source must be (2, ., ., .), got {tuple(source.shape)}")
  ww = omega_np.shape[0]

  pad_zz = _pad_zz(epsilon.shape[3], pml_widths, use_reduced_precision, domain_zz)
  padding = ((absorption_padding, absorption_padding),
             (absorption_padding, absorption_padding), pad_zz)

  rsin = _ramped_sin(omega_np, source_ramp, source_delay, dt, tt)
  if source.shape[3] == 1:  # z-source carries quadrature components.
    source_waveform = np.stack([rsin.imag, rsin.real], axis=-1)
  else:
    source_waveform = np.pad(rsin.imag[:, None], ((0, 0), (0, 1)))

  interval = _sampling_interval(omega_range[0], omega_range[1], ww, dt)
  output_steps = (tt - 2 * interval * ww - 1, tt, interval)
  if output_steps[0] < 0:
    raise ValueError(f"tt={tt} too short for {2 * ww + 1} snapshots at interval {interval}")

  source = _source(source, source_pos, epsilon)
  source_pos = int(source_pos)
  for i in range(3):
    if source.shape[i + 1] == 1:
      source_pos += padding[i][0]
    else:
      pad = [0, 0] * 3
      pad[2 * (2 - i)], pad[2 * (2 - i) + 1] = padding[i]
      source = torch.nn.functional.pad(source, pad)

  if (source.shape[1] == 1 or source.shape[2] == 1) and source_pos % 2 == 1:
    source_pos += 1
    source_waveform = source_waveform[:, ::-1]
  elif source.shape[3] == 1:
    if source.is_complex():
      source = torch.stack([source.imag, source.real])
    else:
      source = torch.stack([torch.zeros_like(source), source])
  if source.is_complex():
    raise ValueError("x/y plane sources must be real")

  xx, yy = epsilon.shape[1], epsilon.shape[2]
  absorption_mask = _absorption_mask(xx + 2 * absorption_padding,
                                     yy + 2 * absorption_padding,
                                     absorption_padding, absorption_coeff)
  zdom = _zz(pml_widths, use_reduced_precision, domain_zz)
  if use_z_as_batch:
    pml_kappa = np.full((zdom, 2), np.inf, np.float32)
    pml_sigma = np.zeros_like(pml_kappa)
    pml_alpha = np.zeros_like(pml_kappa)
  else:
    pml_sigma = _pml_sigma(pml_widths, zdom, pml_sigma_lnr, pml_sigma_m)
    pml_kappa = np.ones_like(pml_sigma)
    pml_alpha = (pml_alpha_coeff * np.ones_like(pml_sigma)).astype(np.float32)

  kwargs = dict(
      epsilon=epsilon,
      dt=dt,
      source_field=source.to(torch.float32).contiguous(),
      source_waveform=np.ascontiguousarray(source_waveform, np.float32),
      source_position=source_pos,
      absorption_mask=absorption_mask,
      pml_kappa=pml_kappa,
      pml_sigma=pml_sigma,
      pml_alpha=pml_alpha,
      pml_widths=tuple(pml_widths),
      output_steps=output_steps,
      use_reduced_precision=use_reduced_precision,
      launch_params=launch_params,
      offset=(padding[0][0], padding[1][0], padding[2][0]),
  )
  return kwargs, omega_np, output_steps


def project_snapshots(fields, omega, output_steps, dt):
  """Snapshots (n_out,3,xx,yy,zz) -> complex phasors (ww,3,xx,yy,zz) (:272-279)."""
  ww = np.asarray(omega).reshape(-1).shape[0]
  phases = _output_phases(omega, output_steps, dt)
  pinv = np.linalg.pinv(phases.T)                       # (2ww, n_out)
  fields = _as_tensor(fields, dtype=torch.float32)
  if fields.is_cuda:
    # one pass over the snapshots, complex phasors written interleaved (b200fdtd_project)
    from . import fdtdz_jax
    return fdtdz_jax.project(fields, pinv.astype(np.float32))
  w = torch.from_numpy(pinv.astype(np.float32)).to(fields.device)
  outputs = torch.einsum("ij,j...->i...", w, fields)
  return torch.complex(outputs[:ww], outputs[ww:])


def field(epsilon, source, omega, source_pos, sim_params, *, engine=None,
          fuse_projection=False):
  """Time-harmonic solution of Maxwell's equations; mirror of ``pjz.field``
  (/root/reference/src/pjz/_field.py:171-279).

  Args:
    epsilon: ``(3, xx, yy, zz)`` permittivity (torch tensor or array).
    source: ``(2, 1, yy, zz)``, ``(2, xx, 1, zz)`` or ``(2, xx, yy, 1)`` excitation.
    omega: ``(ww,)`` angular frequencies.
    source_pos: source plane index along the propagation axis.
    sim_params: ``SimParams``.
    engine: callable with the ``fdtdz_jax.fdtdz`` signature; default = the CUDA engine.
    fuse_projection: form the phasors inside the time-stepping kernels (the engine's
      ``output_projection`` extension) instead of dumping 2ww+1 snapshots and projecting them
      afterwards (:272-279); same pinv weights, applied as a running sum.

  Returns:
    ``(ww, 3, xx, yy, zz)`` complex64 torch tensor.
  """
  engine = engine or _default_engine()
  kwargs, omega_np, output_steps = engine_inputs(epsilon, source, omega, source_pos, sim_params)
  if fuse_projection:
    ww = omega_np.shape[0]
    pinv = np.linalg.pinv(_output_phases(omega_np, output_steps, sim_params.dt).T)
    out = _as_tensor(engine(**kwargs, output_projection=pinv.astype(np.float32)),
                     dtype=torch.float32)
    return torch.complex(out[:ww], out[ww:])
  fields = engine(**kwargs)
  return project_snapshots(fields, omega_np, output_steps, sim_params.dt)


def _amplitudes(beta, vals, x):
  """Forward/backward coefficients from two-plane samples (:293-302)."""
  beta = np.asarray(beta, np.float64).reshape(-1)
  a = np.stack([np.exp(-1j * beta[:, None] * x), np.exp(1j * beta[:, None] * x)], axis=1)
  pinv = np.linalg.pinv(a)                               # (ww, 2planes, 2coef) -> (ww, j, i)
  w = torch.from_numpy(pinv.astype(np.complex64)).to(vals.device)
  return torch.einsum("...ji,...j->...i", w, vals.to(torch.complex64))


def _overlap(mode, beta, pos, is_fwd, output):
  """Mode overlap at two planes next to the port -> (ww, 2) in/out amplitudes (:305-338)."""
  beta = np.asarray(beta, np.float64).reshape(-1).copy()
  if is_fwd is None:
    x = np.array([0, 0])
    sample_at = (pos, pos)
    beta *= 0
  elif is_fwd:
    x = np.array([1, 2])
    sample_at = (pos + 1, pos + 2)
  else:
    x = np.array([-2, -1])
    sample_at = (pos - 2, pos - 1)
    beta *= -1
  mode = _as_tensor(mode, output.device)
  axis = _prop_axis(mode)
  vals = torch.stack(
      [torch.sum(mode * _transverse_slice(output, p, axis), dim=(-4, -3, -2, -1))
       for p in sample_at], dim=-1)
  return _amplitudes(beta, vals, x)


def _overlap_geometry(beta, pos, is_fwd):
  """Sample planes, their offsets from the port and the signed beta of ``_overlap`` (:305-321)."""
  beta = np.asarray(beta, np.float64).reshape(-1).copy()
  if is_fwd is None:
    return (pos, pos), np.array([0, 0]), beta * 0
  if is_fwd:
    return (pos + 1, pos + 2), np.array([1, 2]), beta
  return (pos - 2, pos - 1), np.array([-2, -1]), -beta


def _overlaps_fused(fields, modes, betas, pos, is_fwd):
  """``amplitudes`` and ``svals`` of ``_scatter_impl`` with ALL nports^2 two-plane overlaps formed by
  one kernel launch (``b200fdtd_overlaps``) instead of nports^2 x 2 eager slice-multiply-sum chains;
  the 2x2 least-squares fits (``_amplitudes``) stay on the host-built pinv."""
  from . import fdtdz_jax
  dev = fields[0].device
  geo = [_overlap_geometry(b, p, fwd) for b, p, fwd in zip(betas, pos, is_fwd)]
  ms = [_as_tensor(m, dev) for m in modes]
  axes = ["xyz".find(_prop_axis(m)) for m in ms]
  flat = []
  for m in ms:                                         # (ww|1, 2, xx|1, yy|1, zz|1) -> (ww, 2, U, V)
    m = m.reshape((m.shape[0], 2) + tuple(d for d in m.shape[-3:] if d != 1)) \
        if tuple(m.shape[-3:]).count(1) == 1 else m
    flat.append(m)
  vals = fdtdz_jax.overlaps(fields, flat, axes, [g[0] for g in geo])   # (F, M, 2, ww)
  # the 2x2 least-squares fits of `_amplitudes` (:293-302) for all pairs at once: one pinv stack
  # per port on the host, one einsum on the device
  pinvs = []
  for _, x, beta in geo:
    a = np.stack([np.exp(-1j * beta[:, None] * x), np.exp(1j * beta[:, None] * x)], axis=1)
    pinvs.append(np.linalg.pinv(a))                                    # (ww, plane j, coefficient i)
  w = torch.from_numpy(np.stack(pinvs).astype(np.complex64)).to(dev)  # (M, ww, j, i)
  coefs = torch.einsum("mwji,fmjw->fmwi", w, vals)                     # (F, M, ww, 2): in, out
  amplitudes = []
  for i, fwd in enumerate(is_fwd):
    if fwd is None:
      amplitudes.append(torch.ones(vals.shape[-1], dtype=torch.complex64, device=dev))
    else:
      amplitudes.append(coefs[i, i, :, 0])
  svals = [[coefs[f, m, :, 1] / amplitudes[f] for m in range(len(modes))]
           for f in range(len(fields))]
  return amplitudes, svals


def _scatter_impl(epsilon, omega, modes, betas, pos, is_fwd, sim_params, engine=None,
                  group=None, want_grads=True, fuse_projection=False):
  """Mirror of ``_scatter_impl`` (:346-384).  One independent engine run per port; with a
  ``torch.distributed`` process group the ports are dealt round-robin to the ranks (the
  batch axis of SURVEY.md 8(e)) and the phasor fields are all-gathered afterwards."""
  import torch.distributed as dist
  nports = len(modes)
  world, rank = 1, 0
  # a ``collective`` engine (pjz_b200._decomp.decomposed_engine) uses all ranks for EACH call: every
  # rank runs every port and ends up with the whole field, so nothing is dealt or broadcast here
  collective = bool(getattr(engine, "collective", False))
  if not collective and (group is not None or (dist.is_available() and dist.is_initialized())):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
  epsilon = _as_tensor(epsilon, dtype=torch.float32)
  mine = [i for i in range(nports) if i % world == rank]
  local = {}
  for i in mine:
    m = _as_tensor(modes[i], epsilon.device)
    local[i] = field(epsilon, torch.mean(m, dim=0), omega, pos[i], sim_params, engine=engine,
                     fuse_projection=fuse_projection)
  if world > 1:
    fields = []
    for i in range(nports):
      if i % world == rank:
        buf = local[i].contiguous()
      else:
        ww = np.asarray(omega).reshape(-1).shape[0]
        buf = torch.empty((ww,) + tuple(epsilon.shape), dtype=torch.complex64,
                          device=epsilon.device)
      real = torch.view_as_real(buf)
      dist.broadcast(real, src=dist.get_global_rank(group, i % world) if group else i % world,
                     group=group)
      fields.append(torch.view_as_complex(real))
  else:
    fields = [local[i] for i in range(nports)]

  if fields[0].is_cuda and nports <= 16:
    amplitudes, svals = _overlaps_fused(fields, modes, betas, pos, is_fwd)
  else:
    amplitudes = []
    for f, m, b, p, fwd in zip(fields, modes, betas, pos, is_fwd):
      if fwd is None:
        amplitudes.append(torch.ones(np.asarray(b).reshape(-1).shape[0],
                                     dtype=torch.complex64, device=f.device))
      else:
        amplitudes.append(_overlap(m, b, p, fwd, f)[:, 0])
    svals = [[_overlap(m, b, p, fwd, f)[:, 1] / a
              for m, b, p, fwd in zip(modes, betas, pos, is_fwd)]
             for a, f in zip(amplitudes, fields)]
  grads = None
  if want_grads:
    grads = [[fi * fj / a[:, None, None, None, None] for fj in fields]
             for a, fi in zip(amplitudes, fields)]
  return svals, grads, fields, amplitudes


def _scatter_bwd_fused(fields, amplitudes, g):
  """``_scatter_bwd`` without the N^2 volume temporaries: one CUDA pass over the N phasor fields
  (``b200fdtd_adjoint_reduce``).  ``g[i][j]`` = conj(cotangent), ``(ww,)``."""
  from . import fdtdz_jax
  n = len(fields)
  coef = torch.stack([torch.stack([g[i][j] / amplitudes[i] for j in range(n)]) for i in range(n)])
  return fdtdz_jax.adjoint_reduce(fields, coef)


def _scatter_bwd(grad, g):
  """Mirror of ``_scatter_bwd`` (:393-398): dL/d epsilon = sum_ij sum_ww Re(g_ij grads_ij)."""
  total = None
  for gradi, gi in zip(grad, g):
    for gradij, gij in zip(gradi, gi):
      term = torch.sum(torch.real(gij[:, None, None, None, None] * gradij), dim=0)
      total = term if total is None else total + term
  return total


class _Scatter(torch.autograd.Function):
  """``custom_vjp`` of pjz.scatter (:403-442) as a torch autograd function."""

  @staticmethod
  def forward(ctx, epsilon, omega, modes, betas, pos, is_fwd, sim_params, engine, group,
              fuse_projection):
    # On the GPU the backward pass is the fused product-reduce kernel over the saved phasor
    # fields; the N^2 ``grads`` volumes of the reference are only formed on the CPU test path.
    on_gpu = isinstance(epsilon, torch.Tensor) and epsilon.is_cuda
    svals, grads, fields, amplitudes = _scatter_impl(
        epsilon, omega, modes, betas, pos, is_fwd, sim_params, engine, group,
        want_grads=not on_gpu, fuse_projection=fuse_projection)
    ctx.grads, ctx.fields, ctx.amplitudes = grads, fields, amplitudes
    ctx.n = len(modes)
    return tuple(s for row in svals for s in row)

  @staticmethod
  def backward(ctx, *g):
    n = ctx.n
    zero = torch.zeros_like(ctx.amplitudes[0])
    gm = [[torch.conj(g[i * n + j]) if g[i * n + j] is not None else zero for j in range(n)]
          for i in range(n)]
    if ctx.grads is None:
      return (_scatter_bwd_fused(ctx.fields, ctx.amplitudes, gm),) + (None,) * 9
    return (_scatter_bwd(ctx.grads, gm),) + (None,) * 9


def scatter(epsilon, omega, modes, betas, pos, is_fwd, sim_params, *, engine=None,
            group=None, fuse_projection=False):
  """Scattering values between ``modes``; mirror of ``pjz.scatter``
  (/root/reference/src/pjz/_field.py:403-442).  Returns ``svals[i][j]`` nested lists of
  ``(ww,)`` complex tensors; differentiable w.r.t. ``epsilon`` when it requires grad (the
  backward is pjz's reciprocity formula, :380-398 -- no second engine launch).
  """
  eps_t = _as_tensor(epsilon, dtype=torch.float32)
  n = len(modes)
  if eps_t.requires_grad:
    flat = _Scatter.apply(eps_t, omega, modes, betas, pos, is_fwd, sim_params, engine, group,
                          fuse_projection)
    return [[flat[i * n + j] for j in range(n)] for i in range(n)]
  svals, _, _, _ = _scatter_impl(eps_t, omega, modes, betas, pos, is_fwd, sim_params, engine,
                                 group, want_grads=False, fuse_projection=fuse_projection)
  return svals
