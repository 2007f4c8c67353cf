"""Drop-in for the ``fdtdz_jax`` module that pjz imports (/root/reference/src/pjz/_field.py:6).

``fdtdz(...)`` keeps the keyword signature of the one call pjz makes
(/root/reference/src/pjz/_field.py:254-269) and forwards it to ``libb200fdtd.so`` through
the C ABI declared in ``include/b200fdtd.h``.  There is NO CPU fallback: without the built
library or without a CUDA device the call raises.

Array arguments may be torch tensors (CUDA tensors are used in place, zero-copy, on torch's
current stream and the result is a CUDA tensor) or anything ``np.asarray`` accepts (host path:
``b200fdtd_run_host`` copies up, runs, copies the snapshots back; the result is a NumPy array).

``launch_params`` (opaque in pjz, ``SimParams.launch_params`` :53) may be ``None`` or a dict with
any of ``kernel`` ("auto" | "twopass" | "systolic" | "systolic_async"),
``tile_y``, ``stages``,
``threads``, ``prefetch``.
To make ``import fdtdz_jax`` resolve to this module: ``pjz_b200.fdtdz_jax.install()``.
"""

from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200fdtd.so")

NUM_INPUTS = 7
_KERNELS = {"auto": 0, "twopass": 1, "systolic": 2, "systolic_async": 3,
            "systolic_lean": 5}
ABI_VERSION = 1


class Desc(ctypes.Structure):
  """``b200fdtd_desc`` (include/b200fdtd.h)."""
  _fields_ = ([("struct_bytes", ctypes.c_uint32), ("abi_version", ctypes.c_uint32)] +
              [(n, ctypes.c_int32) for n in (
                  "X", "Y", "Z", "xx", "yy", "zz", "off_x", "off_y", "off_z", "tt",
                  "source_axis", "source_position", "pml_lo", "pml_hi", "out_start",
                  "out_stop", "out_step", "use_reduced_precision")] +
              [("dt", ctypes.c_float)] +
              [(n, ctypes.c_int32) for n in ("kernel", "tile_y", "stages", "threads",
                                             "prefetch", "cols")] +
              [("proj_rows", ctypes.c_int32), ("reserved", ctypes.c_int32)])


_lib = None


def lib():
  """The loaded C-ABI library; raises if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
          f"g.build()'` (there is no CPU fallback for the FDTD engine).")
    L = ctypes.CDLL(LIB_PATH)
    L.b200fdtd_abi_version.restype = ctypes.c_int
    L.b200fdtd_last_error.restype = ctypes.c_char_p
    L.b200fdtd_validate.restype = ctypes.c_int
    L.b200fdtd_num_outputs.restype = ctypes.c_int
    L.b200fdtd_output_bytes.restype = ctypes.c_size_t
    L.b200fdtd_workspace_bytes.restype = ctypes.c_size_t
    L.b200fdtd_run.restype = ctypes.c_int
    L.b200fdtd_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.b200fdtd_run_host.restype = ctypes.c_int
    L.b200fdtd_run_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_int]
    L.b200fdtd_plan_info.restype = ctypes.c_int
    if L.b200fdtd_abi_version() != ABI_VERSION:
      raise RuntimeError("libb200fdtd.so ABI version mismatch")
    _lib = L
  return _lib


def _last_error():
  return lib().b200fdtd_last_error().decode("utf-8", "replace")


def _is_torch(a):
  return type(a).__module__.split(".")[0] == "torch"


def _adopt(a):
  """Foreign device arrays (JAX, CuPy, ... -- anything with ``__dlpack__`` that is neither a torch
  tensor nor a NumPy array) enter zero-copy through DLPack; everything else is returned as is.
  Returns ``(array, adopted)``."""
  if a is None or _is_torch(a) or isinstance(a, np.ndarray) or not hasattr(a, "__dlpack__"):
    return a, False
  import torch
  try:
    return torch.from_dlpack(a), True
  except Exception:                                          # noqa: BLE001
    return a, False                                          # the array interface (host path) still works


def _shape(a):
  return tuple(int(s) for s in a.shape)


def make_desc(epsilon, dt, source_field, source_waveform, source_position, absorption_mask,
              pml_kappa, pml_sigma, pml_alpha, pml_widths, output_steps,
              use_reduced_precision, launch_params, offset, output_projection=None):
  """Shape-checks the arguments the way the reference wrapper would and fills the descriptor."""
  es, ss, ws, ms = (_shape(a) for a in (epsilon, source_field, source_waveform,
                                         absorption_mask))
  ks, gs, als = (_shape(a) for a in (pml_kappa, pml_sigma, pml_alpha))
  if len(es) != 4 or es[0] != 3:
    raise ValueError(f"epsilon must have shape (3, xx, yy, zz), got {es}")
  if len(ms) != 3 or ms[0] != 3:
    raise ValueError(f"absorption_mask must have shape (3, X, Y), got {ms}")
  if len(ks) != 2 or ks[1] != 2 or gs != ks or als != ks:
    raise ValueError(f"pml_kappa/sigma/alpha must share shape (Z, 2), got {ks}, {gs}, {als}")
  if len(ws) != 2 or ws[1] != 2:
    raise ValueError(f"source_waveform must have shape (tt, 2), got {ws}")
  X, Y, Z = ms[1], ms[2], ks[0]
  if len(ss) == 5 and ss == (2, 2, X, Y, 1):
    axis = 2
  elif len(ss) == 4 and ss == (2, 1, Y, Z):
    axis = 0
  elif len(ss) == 4 and ss == (2, X, 1, Z):
    axis = 1
  else:
    raise ValueError(
        f"source_field must have shape (2, 1, {Y}, {Z}), (2, {X}, 1, {Z}) or "
        f"(2, 2, {X}, {Y}, 1), got {ss}")
  if len(pml_widths) != 2 or len(output_steps) != 3 or len(offset) != 3:
    raise ValueError("pml_widths, output_steps, offset must have 2, 3, 3 entries")
  d = Desc()
  d.struct_bytes, d.abi_version = ctypes.sizeof(Desc), ABI_VERSION
  d.X, d.Y, d.Z = X, Y, Z
  d.xx, d.yy, d.zz = es[1:]
  d.off_x, d.off_y, d.off_z = (int(o) for o in offset)
  d.tt = ws[0]
  d.source_axis, d.source_position = axis, int(source_position)
  d.pml_lo, d.pml_hi = int(pml_widths[0]), int(pml_widths[1])
  d.out_start, d.out_stop, d.out_step = (int(v) for v in output_steps)
  d.use_reduced_precision = int(bool(use_reduced_precision))
  d.dt = float(dt)
  lp = launch_params or {}
  if not isinstance(lp, dict):
    raise ValueError("launch_params must be None or a dict (kernel, tile_y, stages, threads, "
                     "prefetch)")
  unknown = set(lp) - {"kernel", "tile_y", "stages", "threads", "prefetch", "cols"}
  if unknown:
    raise ValueError(f"unknown launch_params keys {sorted(unknown)}")
  k = lp.get("kernel", "auto")
  if k not in _KERNELS:
    raise ValueError(f"launch_params['kernel'] must be one of {sorted(_KERNELS)}, got {k!r}")
  d.kernel = _KERNELS[k]
  d.tile_y, d.stages, d.threads, d.prefetch, d.cols = (
      int(lp.get(n, 0)) for n in ("tile_y", "stages", "threads", "prefetch", "cols"))
  if output_projection is not None:
    ps = _shape(output_projection)
    nout = len(range(*d_output_steps(d)))
    if len(ps) != 2 or ps[1] != nout or ps[0] < 1:
      raise ValueError(f"output_projection must have shape (rows, {nout}), got {ps}")
    d.proj_rows = ps[0]
  rc = lib().b200fdtd_validate(ctypes.byref(d))
  if rc != 0:
    raise ValueError(_last_error())
  return d


def d_output_steps(d):
  return (d.out_start, d.out_stop, d.out_step)


def _void_array(ptrs):
  arr = (ctypes.c_void_p * len(ptrs))()
  for i, p in enumerate(ptrs):
    arr[i] = p
  return arr


def fdtdz(epsilon, dt, source_field, source_waveform, source_position, absorption_mask,
          pml_kappa, pml_sigma, pml_alpha, pml_widths, output_steps, use_reduced_precision,
          launch_params=None, offset=(0, 0, 0), *, output_projection=None):
  """Execute an FDTD simulation; signature of ``fdtdz_jax.fdtdz`` as called at
  /root/reference/src/pjz/_field.py:254-269.  Returns ``(n_out, 3, xx, yy, zz)`` float32 E
  snapshots for the steps ``range(*output_steps)``.

  Extension (keyword-only, default off): ``output_projection`` = a ``(rows, n_out)`` matrix W
  fuses the frequency projection of /root/reference/src/pjz/_field.py:272-279 into the time
  stepping: the result is ``(rows, 3, xx, yy, zz)`` with ``out[r] = sum_s W[r, s] * snapshot_s``
  and no snapshot is written."""
  foreign = None                                             # e.g. "jax": results go back the same way
  adopted = []
  for a in (epsilon, source_field, source_waveform, absorption_mask, pml_kappa, pml_sigma,
            pml_alpha, output_projection):
    t, was = _adopt(a)
    if was and foreign is None:
      foreign = type(a).__module__.split(".")[0]
    adopted.append(t)
  (epsilon, source_field, source_waveform, absorption_mask, pml_kappa, pml_sigma, pml_alpha,
   output_projection) = adopted
  if foreign is not None:
    return _return_as(foreign, fdtdz(
        epsilon, dt, source_field, source_waveform, source_position, absorption_mask, pml_kappa,
        pml_sigma, pml_alpha, pml_widths, output_steps, use_reduced_precision, launch_params,
        offset, output_projection=output_projection))
  arrays = [epsilon, source_field, source_waveform, absorption_mask, pml_kappa, pml_sigma,
            pml_alpha]
  if output_projection is not None:
    arrays.append(output_projection)
  d = make_desc(epsilon, dt, source_field, source_waveform, source_position, absorption_mask,
                pml_kappa, pml_sigma, pml_alpha, pml_widths, output_steps,
                use_reduced_precision, launch_params, offset, output_projection)
  L = lib()
  nout = d.proj_rows if d.proj_rows > 0 else L.b200fdtd_num_outputs(ctypes.byref(d))
  out_shape = (nout, 3, d.xx, d.yy, d.zz)
  cuda_dev = None
  for a in arrays:
    if _is_torch(a) and a.is_cuda:
      cuda_dev = a.device
      break

  if cuda_dev is None:
    # Host path (NumPy in, NumPy out).
    host = [np.ascontiguousarray(a.detach().cpu().numpy() if _is_torch(a) else np.asarray(a),
                                 dtype=np.float32) for a in arrays]
    import torch  # device selection + page-locked result buffer
    if not torch.cuda.is_available():
      raise RuntimeError("fdtdz needs a CUDA device (no CPU fallback)")
    # the snapshots come back with one device-to-host copy: land it in page-locked memory
    # (a pageable destination is staged through a bounce buffer at a fraction of the bandwidth)
    out = torch.empty(out_shape, dtype=torch.float32, pin_memory=True).numpy()
    rc = L.b200fdtd_run_host(ctypes.byref(d), _void_array([h.ctypes.data for h in host]),
                             _void_array([out.ctypes.data]), torch.cuda.current_device())
    if rc != 0:
      raise RuntimeError(f"b200fdtd_run_host failed ({rc}): {_last_error()}")
    return out

  import torch
  with torch.cuda.device(cuda_dev):
    dev = [torch.as_tensor(a, dtype=torch.float32, device=cuda_dev).contiguous()
           if _is_torch(a) else
           torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda_dev)
           for a in arrays]
    ws_bytes = L.b200fdtd_workspace_bytes(ctypes.byref(d))
    if ws_bytes == 0:
      raise RuntimeError(f"b200fdtd_workspace_bytes failed: {_last_error()}")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda_dev)
    out = torch.empty(out_shape, dtype=torch.float32, device=cuda_dev)
    stream = torch.cuda.current_stream(cuda_dev).cuda_stream
    rc = L.b200fdtd_run(ctypes.byref(d), _void_array([t.data_ptr() for t in dev]),
                        _void_array([out.data_ptr()]), ws.data_ptr(), ws_bytes, stream)
    if rc != 0:
      raise RuntimeError(f"b200fdtd_run failed ({rc}): {_last_error()}")
    # ws/dev are released to torch's caching allocator, which is stream-ordered.
    return out


def _return_as(foreign, out):
  """Hand the result back to the framework the DLPack inputs came from (JAX: ``jax.dlpack``,
  zero-copy); anything else gets the torch tensor / NumPy array, both of which speak DLPack."""
  if foreign in ("jax", "jaxlib"):
    try:
      import jax.dlpack
      import torch
      return jax.dlpack.from_dlpack(out if _is_torch(out) else torch.from_numpy(out))
    except ImportError:
      pass
  return out


def adjoint_reduce(fields, coef):
  """``out[c,x,y,z] = sum_ij sum_w Re(coef[i,j,w] * fields[i][w] * fields[j][w])`` in one pass
  (``b200fdtd_adjoint_reduce``; replaces the N^2 volume temporaries of
  /root/reference/src/pjz/_field.py:380-398).  ``fields``: N CUDA complex64 tensors
  ``(ww, 3, xx, yy, zz)``; ``coef``: complex ``(N, N, ww)``.  Returns float32 ``(3, xx, yy, zz)``."""
  import torch
  n = len(fields)
  f0 = fields[0]
  if not (_is_torch(f0) and f0.is_cuda):
    raise RuntimeError("adjoint_reduce needs CUDA tensors (no CPU fallback)")
  dev = f0.device
  fs = [f.to(torch.complex64).contiguous() for f in fields]
  ww = int(f0.shape[0])
  nvox = int(f0[0].numel())
  for f in fs:
    if tuple(f.shape) != tuple(f0.shape) or f.device != dev:
      raise ValueError("all phasor fields must share shape and device")
  c = torch.as_tensor(coef, device=dev).to(torch.complex64).contiguous()
  if tuple(c.shape) != (n, n, ww):
    raise ValueError(f"coef must have shape ({n}, {n}, {ww}), got {tuple(c.shape)}")
  out = torch.empty(tuple(f0.shape[1:]), dtype=torch.float32, device=dev)
  L = lib()
  L.b200fdtd_adjoint_reduce.restype = ctypes.c_int
  L.b200fdtd_adjoint_reduce.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
  with torch.cuda.device(dev):
    rc = L.b200fdtd_adjoint_reduce(n, ww, nvox, _void_array([f.data_ptr() for f in fs]),
                                   c.data_ptr(), out.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream)
  if rc != 0:
    raise RuntimeError(f"b200fdtd_adjoint_reduce failed ({rc}): {_last_error()}")
  return out


def project(snapshots, weights):
  """Phasors from snapshots in one pass (``b200fdtd_project``): ``snapshots`` CUDA float32
  ``(n_out, ...)``, ``weights`` ``(2*ww, n_out)`` = the pseudo-inverse of the sampled phases
  (/root/reference/src/pjz/_field.py:272-279).  Returns complex64 ``(ww, ...)``."""
  import torch
  if not (_is_torch(snapshots) and snapshots.is_cuda):
    raise RuntimeError("project needs a CUDA tensor (no CPU fallback)")
  dev = snapshots.device
  x = snapshots.to(torch.float32).contiguous()
  w = torch.as_tensor(np.ascontiguousarray(np.asarray(weights, np.float32))).to(dev)
  n_out = int(x.shape[0])
  if w.ndim != 2 or w.shape[1] != n_out or w.shape[0] % 2:
    raise ValueError(f"weights must have shape (2*ww, {n_out}), got {tuple(w.shape)}")
  ww = int(w.shape[0]) // 2
  nvox = int(x[0].numel())
  out = torch.empty((ww,) + tuple(x.shape[1:]), dtype=torch.complex64, device=dev)
  L = lib()
  L.b200fdtd_project.restype = ctypes.c_int
  L.b200fdtd_project.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
  with torch.cuda.device(dev):
    rc = L.b200fdtd_project(ww, n_out, nvox, x.data_ptr(), w.data_ptr(), out.data_ptr(),
                            torch.cuda.current_stream(dev).cuda_stream)
  if rc != 0:
    raise RuntimeError(f"b200fdtd_project failed ({rc}): {_last_error()}")
  return out


def overlaps(fields, modes, axes, planes):
  """All two-plane mode overlaps at once (``b200fdtd_overlaps``,
  /root/reference/src/pjz/_field.py:305-338).  ``fields``: F CUDA complex64 tensors
  ``(ww, 3, xx, yy, zz)``; ``modes``: M tensors ``(ww, 2, ., ., .)`` with a singleton along the
  port's axis; ``axes``: M ints; ``planes``: ``(M, 2)`` sample planes.  Returns complex64
  ``(F, M, 2, ww)``: ``vals[f, m, k, w] = sum(mode_m[w] * transverse(fields_f[w]) at plane k)``."""
  import torch
  f0 = fields[0]
  if not (_is_torch(f0) and f0.is_cuda):
    raise RuntimeError("overlaps needs CUDA tensors (no CPU fallback)")
  dev = f0.device
  fs = [f.to(torch.complex64).contiguous() for f in fields]
  ww, _, xx, yy, zz = (int(v) for v in f0.shape)
  ms = []
  for m in modes:
    m = torch.as_tensor(m).to(dev)
    if m.shape[0] != ww:
      m = m.expand((ww,) + tuple(m.shape[1:]))
    ms.append(m.to(torch.complex64).contiguous())
  nf, nm = len(fs), len(ms)
  ax = (ctypes.c_int * nm)(*[int(a) for a in axes])
  pl = (ctypes.c_int * (2 * nm))(*[int(p) for row in planes for p in row])
  vals = torch.empty((nf, nm, 2, ww), dtype=torch.complex64, device=dev)
  L = lib()
  L.b200fdtd_overlaps.restype = ctypes.c_int
  L.b200fdtd_overlaps.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p] * 6
  with torch.cuda.device(dev):
    rc = L.b200fdtd_overlaps(nf, nm, ww, xx, yy, zz, _void_array([f.data_ptr() for f in fs]),
                             _void_array([m.data_ptr() for m in ms]), ax, pl, vals.data_ptr(),
                             torch.cuda.current_stream(dev).cuda_stream)
  if rc != 0:
    raise RuntimeError(f"b200fdtd_overlaps failed ({rc}): {_last_error()}")
  return vals


def plan_info(**kwargs):
  """What the engine would launch for these arguments (kernel, tiling, CTAs ...)."""
  d = make_desc(**kwargs)
  info = (ctypes.c_int64 * 8)()
  rc = lib().b200fdtd_plan_info(ctypes.byref(d), info)
  if rc != 0:
    raise RuntimeError(_last_error())
  names = ("kernel", "tile_y", "stages", "threads", "ctas", "smem_bytes", "prefetch",
           "l2_window_mib")
  out = dict(zip(names, (int(v) for v in info)))
  out["kernel"] = {v: k for k, v in _KERNELS.items()}[out["kernel"]]
  return out


def install():
  """Register this module as ``fdtdz_jax`` so that pjz's ``import fdtdz_jax`` binds to it."""
  sys.modules["fdtdz_jax"] = sys.modules[__name__]
