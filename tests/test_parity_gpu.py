"""Parity tests proper: the CUDA engine, called through the C ABI (``pjz_b200.fdtdz_jax`` ->
``libb200fdtd.so``), against the oracle on identical seeded inputs.

Bar (BASELINE.json north_star): relative L2 <= 1e-5 in fp32 mode against the float64 spec.  We
also demand BIT-EXACT equality with the C fp32 oracle, whose per-cell operation order the
kernels share (oracle/fdtd_c.c, pjz_b200/csrc/fdtd_common.cuh).  Reduced precision (fp16
storage): bit-exact against the C oracle's fp16-storage mode and rel-L2 <= 5e-3 against the
fp32 run on these short runs (the stated bound of the mode, 3e-2 for runs of up to 20 000 steps, is
measured in tests/test_parity_configs_gpu.py).  At BASELINE sizes, where the oracle is too slow, size-independent properties are
used: the two independent kernels agree bit-for-bit, linearity in the source, schedule
selection.
"""

import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fdtd_c, fdtd_numpy
from pjz_b200 import fdtdz_jax
from tests.golden.make_golden import CASES
from tests.problems import random_problem, rel_l2

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "engine_golden.npz"))
KERNELS = ["twopass", "systolic", "systolic_async"]
FP32_TOL = 1e-5


def run_gpu(kw, **launch):
  kw = dict(kw)
  kw["launch_params"] = launch or None
  kw["epsilon"] = torch.from_numpy(np.ascontiguousarray(kw["epsilon"])).cuda()
  out = fdtdz_jax.fdtdz(**kw)
  assert out.is_cuda and out.dtype == torch.float32
  return out.cpu().numpy()


@pytest.fixture(scope="module", autouse=True)
def _built(built):
  assert torch.cuda.is_available()
  assert os.path.exists(fdtdz_jax.LIB_PATH)


@pytest.mark.parametrize("kernel", KERNELS + ["auto"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_vectors(name, kernel):
  kw = random_problem(**CASES[name])
  out = run_gpu(kw, kernel=kernel)
  assert out.shape == GOLDEN[name].shape
  assert rel_l2(out, GOLDEN[name]) <= FP32_TOL
  np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("name", ["tall_128", "short_64", "short_30"])
def test_golden_vectors_lean_kernels(name):
  """The committed golden vectors whose column height fits the warp-per-column-pair kernel
  (32 fp32 vectors) and its sub-warp variants (16 and 8 vectors), through kernel="systolic_lean"."""
  kw = random_problem(**CASES[name])
  out = run_gpu(kw, kernel="systolic_lean")
  assert out.shape == GOLDEN[name].shape
  assert rel_l2(out, GOLDEN[name]) <= FP32_TOL
  np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("pml", [(0, 0), (4, 6)])
def test_stress_all_orientations(axis, pml, kernel):
  # SURVEY.md 8(d) "stress": 48x40x32, eps ~ U[1, 12.25], random source, every orientation
  kw = random_problem(domain=(48, 40, 32), sub=(40, 30, 20), axis=axis, pml=pml, tt=60,
                      seed=2 + axis, output_steps=(30, 60, 7))
  out = run_gpu(kw, kernel=kernel)
  assert rel_l2(out, fdtd_numpy.fdtdz(**kw)) <= FP32_TOL
  np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("domain,pml", [
    ((20, 18, 96), (16, 16)),    # pjz's default reduced-precision height: Zq = 24 straddles warps
    ((10, 12, 128), (16, 16)),   # one warp per z-column
    ((9, 7, 13), (3, 5)),        # ragged z (padding lanes), odd everything
    ((6, 40, 8), (2, 2)),        # several columns per warp
    ((5, 6, 256), (10, 12)),     # z-column spans two warps
    ((1, 9, 8), (0, 3)), ((7, 1, 8), (2, 0)), ((6, 5, 1), (0, 0)), ((2, 2, 4), (1, 1)),
])
def test_ragged_and_degenerate_domains(domain, pml, kernel):
  for axis in (0, 2):
    kw = random_problem(domain=domain, sub=domain, offset=(0, 0, 0), axis=axis, pml=pml, tt=14,
                        seed=7, output_steps=(5, 14, 4), absorb_pad=min(2, min(domain[:2]) // 2))
    out = run_gpu(kw, kernel=kernel)
    np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("cols", [1, 2])
@pytest.mark.parametrize("prefetch", [1, 2, 3])
@pytest.mark.parametrize("tile_y,stages", [(1, 2), (3, 5), (6, 3), (4, 40)])
def test_systolic_async_tilings(tile_y, stages, prefetch, cols):
  kw = random_problem(domain=(11, 23, 16), axis=1, pml=(4, 4), tt=45, seed=13,
                      output_steps=(20, 45, 6))
  want = fdtd_c.fdtdz(**kw)
  out = run_gpu(kw, kernel="systolic_async", tile_y=tile_y, stages=stages, prefetch=prefetch,
                cols=cols)
  np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("domain,pml", [((20, 18, 96), (16, 16)), ((10, 12, 128), (16, 16)),
                                        ((9, 7, 13), (3, 5)), ((5, 6, 256), (10, 12))])
def test_systolic_async_one_column_per_thread(domain, pml):
  for axis in (0, 2):
    kw = random_problem(domain=domain, sub=domain, offset=(0, 0, 0), axis=axis, pml=pml, tt=14,
                        seed=7, output_steps=(5, 14, 4), absorb_pad=2)
    np.testing.assert_array_equal(run_gpu(kw, kernel="systolic_async", cols=1),
                                  fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("tile_y,stages", [(0, 0), (1, 2), (2, 3), (3, 5), (6, 3), (4, 40), (13, 2)])
def test_systolic_lean_tilings(tile_y, stages, axis):
  """The warp-per-column-pair kernel (fp32, z-column of exactly 32 vectors): every tiling, all
  source orientations, bit-exact against the C oracle."""
  kw = random_problem(domain=(11, 23, 128), axis=axis, pml=(16, 16), tt=45, seed=13,
                      output_steps=(20, 45, 6))
  want = fdtd_c.fdtdz(**kw)
  out = run_gpu(kw, kernel="systolic_lean", tile_y=tile_y, stages=stages)
  np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("domain,pml,zb", [((9, 14, 125), (3, 5), False), ((16, 1, 128), (0, 0), False),
                                           ((5, 2, 127), (10, 12), False), ((12, 26, 128), (0, 0), True),
                                           ((3, 30, 126), (16, 16), False), ((2, 5, 128), (4, 4), False)])
def test_systolic_lean_ragged_domains(domain, pml, zb):
  for axis in (0, 1, 2):
    kw = random_problem(domain=domain, sub=domain, offset=(0, 0, 0), axis=axis, pml=pml, tt=14,
                        seed=7, output_steps=(5, 14, 4), absorb_pad=2, z_as_batch=zb)
    np.testing.assert_array_equal(run_gpu(kw, kernel="systolic_lean"), fdtd_c.fdtdz(**kw))


def test_systolic_lean_rejects_other_geometries():
  # neither 32 fp32 vectors per column (warp per column pair) nor <= 16 (half-warp variant)
  kw = random_problem(domain=(8, 8, 72), tt=4, seed=1)
  with pytest.raises((ValueError, RuntimeError)):
    run_gpu(kw, kernel="systolic_lean")
  kw = random_problem(domain=(8, 8, 136), tt=4, seed=1, reduced=True)
  with pytest.raises((ValueError, RuntimeError)):
    run_gpu(kw, kernel="systolic_lean")


# ---- half-warp variant of the lean kernel (kernels_lean16.cuh): columns of <= 16 vectors ---------

@pytest.mark.parametrize("reduced,Z", [(True, 96), (True, 128), (False, 64), (False, 30)])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("tile_y,stages", [(0, 0), (1, 2), (2, 3), (3, 5), (6, 3), (4, 40), (21, 2), (31, 2)])
def test_systolic_lean_halfwarp_tilings(tile_y, stages, axis, reduced, Z):
  """pjz's default engine geometry (fp16 storage, 96 z-cells) and its neighbours: every tiling,
  all source orientations, bit-exact against the C oracle (fp16 storage mode included)."""
  kw = random_problem(domain=(11, 23, Z), axis=axis, pml=(16, 16) if Z >= 64 else (5, 7), tt=45,
                      seed=13, output_steps=(20, 45, 6), reduced=reduced)
  want = fdtd_c.fdtdz(**kw)
  out = run_gpu(kw, kernel="systolic_lean", tile_y=tile_y, stages=stages)
  np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("reduced,domain,pml,zb", [
    (True, (9, 14, 121), (3, 5), False), (True, (16, 1, 128), (0, 0), False),
    (True, (5, 2, 127), (10, 12), False), (True, (12, 26, 96), (0, 0), True),
    (True, (3, 30, 64), (16, 16), False), (True, (7, 9, 13), (3, 5), False),
    (True, (6, 5, 1), (0, 0), False), (True, (2, 5, 8), (4, 4), False),
    (False, (9, 14, 61), (3, 5), False), (False, (16, 1, 64), (0, 0), False),
    (False, (5, 2, 63), (10, 12), False), (False, (12, 26, 32), (0, 0), True),
    (False, (3, 30, 48), (16, 16), False), (False, (7, 9, 13), (3, 5), False),
    (False, (6, 5, 1), (0, 0), False), (False, (2, 9, 8), (0, 3), False),
    (False, (20, 47, 40), (5, 6), False), (True, (20, 47, 100), (12, 9), False)])
def test_systolic_lean_halfwarp_ragged_domains(reduced, domain, pml, zb):
  for axis in (0, 1, 2):
    kw = random_problem(domain=domain, sub=domain, offset=(0, 0, 0), axis=axis, pml=pml, tt=14,
                        seed=7, output_steps=(5, 14, 4), absorb_pad=min(2, min(domain[:2]) // 2),
                        z_as_batch=zb, reduced=reduced)
    np.testing.assert_array_equal(run_gpu(kw, kernel="systolic_lean"), fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("reduced,Z", [(True, 96), (True, 60), (False, 64), (False, 32), (False, 14)])
@pytest.mark.parametrize("tile_y,stages", [(0, 0), (5, 3), (9, 2), (17, 2), (40, 2), (63, 1)])
def test_systolic_lean_subwarp_wide_tiles(tile_y, stages, reduced, Z):
  """Many columns per warp (8 lanes per column and / or 2 columns per thread): tiles wider and
  narrower than a warp's column group, tiles that end inside one."""
  for axis in (1, 2):
    kw = random_problem(domain=(9, 70, Z), axis=axis, pml=(5, 7), tt=21, seed=23 + axis,
                        output_steps=(9, 21, 5), reduced=reduced)
    out = run_gpu(kw, kernel="systolic_lean", tile_y=tile_y, stages=stages)
    np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("reduced,Z", [(True, 96), (False, 64)])
def test_systolic_lean_halfwarp_fused_projection(reduced, Z):
  kw = random_problem(domain=(9, 21, Z), axis=0, pml=(8, 8), tt=33, seed=91, output_steps=(8, 33, 4),
                      reduced=reduced)
  W = np.random.default_rng(7).standard_normal((4, len(range(8, 33, 4)))).astype(np.float32)
  want = fdtd_c.fdtdz(**kw, output_projection=W)
  dev = dict(kw)
  dev["launch_params"] = {"kernel": "systolic_lean"}
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  np.testing.assert_array_equal(fdtdz_jax.fdtdz(**dev, output_projection=W).cpu().numpy(), want)


@pytest.mark.parametrize("reduced,Z", [(True, 96), (True, 128), (True, 64), (False, 64), (False, 32)])
def test_systolic_lean_halfwarp_long_run_at_pjz_default_size(reduced, Z):
  """256x256 in x-y at pjz's default height, 300 steps through the full-size plan (L2 discard on
  when a column is whole 128-byte lines): bit-for-bit against the cp.async kernel, which shares
  neither the staging nor the discard logic; reduced precision within the stated bound of fp32."""
  kw = random_problem(domain=(256, 256, Z), sub=(192, 192, Z - 20), offset=(32, 32, 10), axis=0,
                      pml=(16, 16) if Z > 32 else (8, 8), tt=300, seed=6, output_steps=(120, 300, 89), absorb_pad=32,
                      absorb_coeff=1e-4, reduced=reduced)
  a = run_gpu(kw, kernel="systolic_async")
  b = run_gpu(kw, kernel="systolic_lean")
  np.testing.assert_array_equal(a, b)
  assert np.isfinite(a).all() and np.abs(a[-1]).max() > 0


@pytest.mark.parametrize("tile_y,stages", [(1, 2), (3, 5), (5, 3), (14, 7), (4, 40)])
def test_systolic_tilings(tile_y, stages):
  """Every tiling / pipeline depth must give the same bits (many stages on a short x extent
  wraps the sweep window around the periodic boundary)."""
  kw = random_problem(domain=(11, 23, 16), axis=1, pml=(4, 4), tt=45, seed=13,
                      output_steps=(20, 45, 6))
  want = fdtd_c.fdtdz(**kw)
  out = run_gpu(kw, kernel="systolic", tile_y=tile_y, stages=stages)
  np.testing.assert_array_equal(out, want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_z_as_batch(kernel):
  kw = random_problem(domain=(16, 12, 8), axis=0, pml=(0, 0), tt=30, seed=8, z_as_batch=True,
                      output_steps=(10, 30, 5))
  out = run_gpu(kw, kernel=kernel)
  np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("axis", [0, 2])
def test_reduced_precision(axis, kernel):
  kw = random_problem(domain=(24, 20, 40), axis=axis, pml=(5, 6), tt=50, seed=31, reduced=True,
                      output_steps=(25, 50, 6))
  out = run_gpu(kw, kernel=kernel)
  np.testing.assert_array_equal(out, fdtd_c.fdtdz(**kw))
  full = fdtd_numpy.fdtdz(**{**kw, "use_reduced_precision": False})
  assert rel_l2(out, full) <= 5e-3     # 50 steps; long runs: tests/test_parity_configs_gpu.py


def test_host_path_equals_device_path():
  kw = random_problem(domain=(20, 16, 24), axis=0, tt=25, seed=17, output_steps=(5, 25, 5))
  dev = run_gpu(kw)
  host = fdtdz_jax.fdtdz(**kw)           # NumPy in -> b200fdtd_run_host -> NumPy out
  assert isinstance(host, np.ndarray)
  np.testing.assert_array_equal(host, dev)


def test_auto_plan_prefers_systolic_and_reports():
  kw = random_problem(domain=(64, 64, 48), tt=20, seed=1)
  info = fdtdz_jax.plan_info(**kw)
  assert info["kernel"] == "systolic_async" and info["ctas"] >= 1 and info["prefetch"] >= 1
  # columns of <= 16 vectors whose lanes are (nearly) all busy: the sub-warp lean kernel
  for domain, reduced, want in [((64, 64, 64), False, "systolic_lean"), ((64, 64, 128), True, "systolic_lean"),
                                ((64, 64, 60), True, "systolic_lean"), ((64, 64, 96), True, "systolic_lean"),
                                ((64, 64, 80), True, "systolic_async"),
                                ((64, 64, 32), False, "systolic_async"), ((64, 64, 128), False, "systolic_lean")]:
    kw = random_problem(domain=domain, tt=20, seed=1, reduced=reduced)
    assert fdtdz_jax.plan_info(**kw)["kernel"] == want, (domain, reduced)


def test_no_outputs_and_zero_steps():
  kw = random_problem(domain=(8, 8, 8), tt=6, seed=1, output_steps=(6, 6, 1))
  assert run_gpu(kw).shape[0] == 0


def test_schedule_selection_and_linearity_large():
  """BASELINE cfg2-sized domain (256x256x128, fp32), few steps: independent kernels agree
  bit-for-bit, output selection is a pure subset, the update is linear in the source."""
  rng = np.random.default_rng(5)
  X, Y, Z = 256, 256, 128
  kw = random_problem(domain=(X, Y, Z), sub=(192, 192, 96), offset=(32, 32, 16), axis=0,
                      pml=(16, 16), tt=16, seed=5, output_steps=(0, 16, 5), absorb_pad=32,
                      absorb_coeff=1e-4)
  a = run_gpu(kw, kernel="twopass")
  b = run_gpu(kw, kernel="systolic")
  np.testing.assert_array_equal(a, b)
  np.testing.assert_array_equal(a, run_gpu(kw, kernel="systolic_async"))
  np.testing.assert_array_equal(a, run_gpu(kw, kernel="systolic_lean"))
  assert np.isfinite(a).all() and np.abs(a).max() > 0
  kw2 = dict(kw); kw2["output_steps"] = (5, 16, 10)
  np.testing.assert_array_equal(run_gpu(kw2), a[1:4:2])
  kw3 = dict(kw); kw3["source_field"] = kw["source_field"] * 2
  assert rel_l2(run_gpu(kw3), 2 * a) <= 1e-6
  # a thin slab of the big run against the oracle on the identical inputs would need the whole
  # domain; instead spot-check with the C oracle restricted to 4 steps.
  kw4 = dict(kw); kw4["source_waveform"] = kw["source_waveform"][:4]; kw4["output_steps"] = (3, 4, 1)
  np.testing.assert_array_equal(run_gpu(kw4, kernel="systolic"), fdtd_c.fdtdz(**kw4))


def test_workspace_too_small_is_an_error_not_a_crash():
  kw = random_problem(domain=(16, 16, 16), tt=4, seed=2)
  d = fdtdz_jax.make_desc(**kw)
  L = fdtdz_jax.lib()
  need = L.b200fdtd_workspace_bytes(ctypes.byref(d))
  assert need > 0
  ins = [torch.from_numpy(np.ascontiguousarray(kw[k], np.float32)).cuda() for k in (
      "epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa", "pml_sigma",
      "pml_alpha")]
  out = torch.empty(L.b200fdtd_output_bytes(ctypes.byref(d)) // 4, device="cuda")
  ws = torch.empty(need // 2, dtype=torch.uint8, device="cuda")
  in_arr = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in ins])
  out_arr = (ctypes.c_void_p * 1)(out.data_ptr())
  rc = L.b200fdtd_run(ctypes.byref(d), in_arr, out_arr, ws.data_ptr(), need // 2, None)
  assert rc == 2 and b"workspace too small" in L.b200fdtd_last_error()
  in_arr[3] = None
  rc = L.b200fdtd_run(ctypes.byref(d), in_arr, out_arr, None, 0, None)
  assert rc == 1 and b"inputs[3]" in L.b200fdtd_last_error()


def test_xla_custom_call_entry_and_internal_workspace():
  kw = random_problem(domain=(16, 12, 16), tt=10, seed=3, output_steps=(4, 10, 3))
  want = run_gpu(kw)
  d = fdtdz_jax.make_desc(**kw)
  L = fdtdz_jax.lib()
  ins = [torch.from_numpy(np.ascontiguousarray(kw[k], np.float32)).cuda() for k in (
      "epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa", "pml_sigma",
      "pml_alpha")]
  out = torch.zeros(want.shape, device="cuda")
  ws = torch.empty(L.b200fdtd_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device="cuda")
  bufs = (ctypes.c_void_p * 9)(*([t.data_ptr() for t in ins] + [out.data_ptr(), ws.data_ptr()]))
  opaque = ctypes.string_at(ctypes.byref(d), ctypes.sizeof(d))
  L.b200fdtd_xla_custom_call.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p,
                                         ctypes.c_size_t]
  L.b200fdtd_xla_custom_call(None, bufs, opaque, len(opaque))
  torch.cuda.synchronize()
  np.testing.assert_array_equal(out.cpu().numpy(), want)
  # engine-allocated (stream-ordered) workspace
  out2 = torch.zeros(want.shape, device="cuda")
  rc = L.b200fdtd_run(ctypes.byref(d), (ctypes.c_void_p * 7)(*[t.data_ptr() for t in ins]),
                      (ctypes.c_void_p * 1)(out2.data_ptr()), None, 0, None)
  assert rc == 0
  torch.cuda.synchronize()
  np.testing.assert_array_equal(out2.cpu().numpy(), want)


def test_xla_status_returning_custom_call_reports_failures(tmp_path):
  """b200fdtd_xla_custom_call_status: the result is computed on success; a bad descriptor or a
  scratch buffer smaller than this device's plan needs sets the XLA status (here a stand-in
  `XlaCustomCallStatusSetFailure` loaded into the process, as jaxlib would provide it) instead of
  returning garbage silently."""
  import struct
  import subprocess
  stub_c = tmp_path / "xla_status_stub.c"
  stub_c.write_text(
      "#include <string.h>\n#include <stddef.h>\n"
      "void XlaCustomCallStatusSetFailure(void* s, const char* m, size_t n) {\n"
      "  if (n > 255) n = 255; memcpy(s, m, n); ((char*)s)[n] = 0; }\n")
  stub_so = tmp_path / "libxla_status_stub.so"
  subprocess.check_call(["/usr/bin/gcc", "-shared", "-fPIC", str(stub_c), "-o", str(stub_so)])
  ctypes.CDLL(str(stub_so), mode=ctypes.RTLD_GLOBAL)
  kw = random_problem(domain=(16, 12, 16), tt=10, seed=3, output_steps=(4, 10, 3))
  want = run_gpu(kw)
  d = fdtdz_jax.make_desc(**kw)
  L = fdtdz_jax.lib()
  L.b200fdtd_workspace_bytes.restype = ctypes.c_size_t
  need = L.b200fdtd_workspace_bytes(ctypes.byref(d))
  ins = [torch.from_numpy(np.ascontiguousarray(kw[k], np.float32)).cuda() for k in (
      "epsilon", "source_field", "source_waveform", "absorption_mask", "pml_kappa", "pml_sigma",
      "pml_alpha")]
  out = torch.zeros(want.shape, device="cuda")
  ws = torch.empty(need, dtype=torch.uint8, device="cuda")
  bufs = (ctypes.c_void_p * 9)(*([t.data_ptr() for t in ins] + [out.data_ptr(), ws.data_ptr()]))
  fn = L.b200fdtd_xla_custom_call_status
  fn.restype = None
  fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
  desc_bytes = ctypes.string_at(ctypes.byref(d), ctypes.sizeof(d))
  status = ctypes.create_string_buffer(256)
  opaque = desc_bytes + struct.pack("<Q", need)
  fn(None, bufs, opaque, len(opaque), status)
  torch.cuda.synchronize()
  assert status.value == b""
  np.testing.assert_array_equal(out.cpu().numpy(), want)
  # scratch declared at trace time is smaller than this device's plan needs
  out.zero_()
  opaque = desc_bytes + struct.pack("<Q", need // 2)
  fn(None, bufs, opaque, len(opaque), status)
  torch.cuda.synchronize()
  assert b"scratch declared at trace time" in status.value
  assert not out.any()
  # malformed opaque
  status = ctypes.create_string_buffer(256)
  fn(None, bufs, desc_bytes[:-4], len(desc_bytes) - 4, status)
  assert b"opaque must be a b200fdtd_desc" in status.value


def test_two_streams_are_independent():
  kw1 = random_problem(domain=(24, 20, 16), tt=20, seed=41)
  kw2 = random_problem(domain=(24, 20, 16), tt=20, seed=42)
  w1, w2 = fdtd_c.fdtdz(**kw1), fdtd_c.fdtdz(**kw2)
  s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
  kw1["epsilon"] = torch.from_numpy(kw1["epsilon"]).cuda()
  kw2["epsilon"] = torch.from_numpy(kw2["epsilon"]).cuda()
  torch.cuda.synchronize()
  with torch.cuda.stream(s1):
    o1 = fdtdz_jax.fdtdz(**kw1)
  with torch.cuda.stream(s2):
    o2 = fdtdz_jax.fdtdz(**kw2)
  torch.cuda.synchronize()
  np.testing.assert_array_equal(o1.cpu().numpy(), w1)
  np.testing.assert_array_equal(o2.cpu().numpy(), w2)


def test_field_through_cuda_engine_matches_oracle_engine():
  """pjz.field() glue over the CUDA engine == the same glue over the oracle (cfg1-style)."""
  from pjz_b200 import SimParams, field, mode
  omega = np.array([2 * np.pi / 37])
  eps = np.ones((3, 40, 30, 20), np.float32)
  eps[:, :, 9:21, 8:12] = 12.25
  beta, exc, _, _ = mode(eps[:, 6:7], omega, 1)
  p = SimParams(omega_range=(omega[0], omega[0]), tt=600, dt=0.5, absorption_padding=10,
                absorption_coeff=4e-4, pml_widths=(6, 6), use_reduced_precision=False,
                domain_zz=32)
  want = field(eps, exc[0, ..., 0], omega, 6, p, engine=fdtd_c.fdtdz)
  got = field(torch.from_numpy(eps).cuda(), exc[0, ..., 0], omega, 6, p)
  assert got.is_cuda and got.dtype == torch.complex64
  # snapshots are bit-identical; the pinv projection einsum runs on the GPU vs the CPU
  torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6 * float(want.abs().max()))


# ---- fused frequency projection ---------------------------------------------

@pytest.mark.parametrize("kernel,domain,pml", [
    ("twopass", (12, 10, 16), (3, 4)), ("systolic", (12, 10, 16), (3, 4)),
    ("systolic_async", (14, 11, 24), (4, 4)),
    ("systolic_lean", (9, 21, 128), (16, 16)), ("auto", (7, 30, 126), (4, 6))])
def test_fused_projection_is_bit_exact(kernel, domain, pml):
  """``output_projection``: the kernels accumulate W[:, s] * snapshot_s in place at every output
  step (no snapshot is written); bit-exact against the C oracle's fmaf running sum."""
  for axis in (0, 2):
    kw = random_problem(domain=domain, axis=axis, pml=pml, tt=33, seed=91 + axis,
                        output_steps=(8, 33, 4))
    W = np.random.default_rng(7).standard_normal((4, len(range(8, 33, 4)))).astype(np.float32)
    want = fdtd_c.fdtdz(**kw, output_projection=W)
    dev = dict(kw)
    dev["launch_params"] = {"kernel": kernel}
    dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
    got = fdtdz_jax.fdtdz(**dev, output_projection=W)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    host = dict(kw)
    host["launch_params"] = {"kernel": kernel}
    np.testing.assert_array_equal(fdtdz_jax.fdtdz(**host, output_projection=W), want)  # host path


@pytest.mark.parametrize("kernel,domain", [("systolic_lean", (9, 21, 128)), ("twopass", (10, 9, 32)),
                                           ("systolic_async", (10, 12, 24))])
def test_fused_projection_aligned_crop_takes_the_vector_path(kernel, domain):
  """Crop aligned to 16 bytes in z (offset and height multiples of 4): the accumulators are
  updated with whole float4 read-modify-writes; same bits as the oracle's running sum."""
  X, Y, Z = domain
  kw = random_problem(domain=domain, sub=(X - 2, Y - 3, Z - 8), offset=(1, 2, 4), axis=0, pml=(4, 4),
                      tt=26, seed=17, output_steps=(5, 26, 5))
  W = np.random.default_rng(3).standard_normal((3, len(range(5, 26, 5)))).astype(np.float32)
  want = fdtd_c.fdtdz(**kw, output_projection=W)
  dev = dict(kw)
  dev["launch_params"] = {"kernel": kernel}
  dev["epsilon"] = torch.from_numpy(kw["epsilon"]).cuda()
  np.testing.assert_array_equal(fdtdz_jax.fdtdz(**dev, output_projection=W).cpu().numpy(), want)


def test_fused_projection_rejects_bad_shapes_and_reduced_matches():
  kw = random_problem(domain=(12, 10, 16), tt=20, seed=5, output_steps=(5, 20, 5), reduced=True)
  W = np.ones((2, 3), np.float32)
  got = fdtdz_jax.fdtdz(**kw, output_projection=W)
  np.testing.assert_array_equal(got, fdtd_c.fdtdz(**kw, output_projection=W))
  with pytest.raises(ValueError):
    fdtdz_jax.fdtdz(**kw, output_projection=np.ones((2, 4), np.float32))


def test_field_fused_projection_on_the_gpu():
  from pjz_b200 import SimParams, field
  omega = np.array([2 * np.pi / 37, 2 * np.pi / 33])
  eps = np.ones((3, 40, 30, 20), np.float32)
  eps[:, :, 9:21, 8:12] = 12.25
  src = np.random.default_rng(3).standard_normal((2, 1, 30, 20)).astype(np.float32)
  p = SimParams(omega_range=(omega[0], omega[1]), tt=700, dt=0.5, absorption_padding=10,
                absorption_coeff=4e-4, pml_widths=(6, 6), use_reduced_precision=False,
                domain_zz=32)
  e = torch.from_numpy(eps).cuda()
  want = field(e, src, omega, 6, p)
  got = field(e, src, omega, 6, p, fuse_projection=True)
  torch.testing.assert_close(got, want, rtol=1e-4, atol=2e-6 * float(want.abs().max()))


# ---- adjoint product-reduce ----------------------------------------------------

@pytest.mark.parametrize("nports,ww,shape", [(1, 1, (5, 4, 3)), (2, 3, (9, 8, 7)), (4, 2, (16, 12, 10)),
                                             (8, 4, (20, 16, 12)), (16, 1, (6, 5, 4)),
                                             # > 48 KB of coefficient table in shared memory
                                             (16, 26, (3, 2, 2)), (12, 64, (2, 2, 2))])
def test_adjoint_reduce_matches_the_reference_formula(nports, ww, shape):
  """b200fdtd_adjoint_reduce vs the mirror of pjz's N^2-temporaries formula
  (_field.py:380-398 -> pjz_b200/_field.py:_scatter_bwd), float32 tolerance 1e-5."""
  from pjz_b200 import _field as glue
  g = torch.Generator(device="cpu").manual_seed(nports * 10 + ww)
  fields = [torch.complex(torch.randn((ww, 3) + shape, generator=g),
                          torch.randn((ww, 3) + shape, generator=g)).cuda() for _ in range(nports)]
  amps = [torch.complex(torch.randn(ww, generator=g), torch.randn(ww, generator=g)).cuda() + 2
          for _ in range(nports)]
  gm = [[torch.complex(torch.randn(ww, generator=g), torch.randn(ww, generator=g)).cuda()
         for _ in range(nports)] for _ in range(nports)]
  grads = [[fi * fj / a[:, None, None, None, None] for fj in fields] for a, fi in zip(amps, fields)]
  want = glue._scatter_bwd(grads, gm)
  got = glue._scatter_bwd_fused(fields, amps, gm)
  assert got.shape == want.shape and got.dtype == torch.float32
  assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) <= 1e-5


def test_scatter_gradient_through_the_fused_backward():
  """scatter() on CUDA tensors: forward through the CUDA engine, backward through the fused
  product-reduce kernel; same gradient as the N^2-temporaries formula."""
  from pjz_b200 import SimParams, mode, scatter
  from pjz_b200 import _field as glue
  omega = np.array([2 * np.pi / 37])
  eps = np.ones((3, 40, 30, 20), np.float32)
  eps[:, :, 9:21, 8:12] = 12.25
  b0, x0, _, _ = mode(eps[:, 6:7], omega, 1)
  b1, x1, _, _ = mode(eps[:, 33:34], omega, 1)
  m0, m1 = x0[..., 0], x1[..., 0]
  p = SimParams(omega_range=(omega[0], omega[0]), tt=500, dt=0.5, absorption_padding=10,
                absorption_coeff=4e-4, pml_widths=(6, 6), use_reduced_precision=False,
                domain_zz=32)
  args = (omega, [m0, m1], [b0[:, 0], b1[:, 0]], [6, 33], [True, False], p)
  e = torch.from_numpy(eps).cuda().requires_grad_(True)
  sv = scatter(e, *args, fuse_projection=True)
  loss = sum((s.abs() ** 2).sum() for row in sv for s in row)
  loss.backward()
  assert e.grad is not None and e.grad.shape == e.shape and torch.isfinite(e.grad).all()
  # reference formula on the same fields
  svals, grads, fields, amps = glue._scatter_impl(e.detach(), *args, want_grads=True,
                                                  fuse_projection=True)
  # torch hands backward() g = 2 s for L = |s|^2; _Scatter.backward conjugates it
  want = glue._scatter_bwd(grads, [[torch.conj(2 * s) for s in row] for row in svals])
  assert rel_l2(e.grad.cpu().numpy(), want.cpu().numpy()) <= 1e-4


def test_lean_long_run_matches_independent_kernel_at_baseline_size():
  """cfg2-sized domain, 400 steps (50 pipeline rounds): the default plan -- L2 discard of
  consumed planes and progress-gated prefetch included -- against the cp.async kernel, which
  shares neither the staging nor the discard logic.  Bit-for-bit on three snapshots."""
  kw = random_problem(domain=(256, 256, 128), sub=(192, 192, 96), offset=(32, 32, 16), axis=0,
                      pml=(16, 16), tt=400, seed=6, output_steps=(150, 400, 120), absorb_pad=32,
                      absorb_coeff=1e-4)
  a = run_gpu(kw, kernel="systolic_async")
  b = run_gpu(kw)                       # AUTO -> systolic_lean
  assert fdtdz_jax.plan_info(**{**kw, "launch_params": None})["kernel"] == "systolic_lean"
  np.testing.assert_array_equal(a, b)
  assert np.isfinite(a).all() and np.abs(a[-1]).max() > 0


# ---- one-pass snapshot projection and fused port overlaps ---------------------

@pytest.mark.parametrize("ww,shape", [(1, (3, 8, 6, 4)), (2, (3, 9, 7, 5)), (4, (3, 40, 30, 20)),
                                      (8, (3, 12, 10, 8)), (11, (3, 6, 5, 4)), (3, (3, 5, 3, 3))])
def test_project_kernel_matches_the_reference_einsum(ww, shape):
  """b200fdtd_project vs `einsum("ij,j...->i...", pinv, snapshots)` + complex()
  (/root/reference/src/pjz/_field.py:272-279): <= 1e-6 of the largest phasor."""
  g = torch.Generator().manual_seed(ww)
  n_out = 2 * ww + 1
  snaps = torch.randn((n_out,) + shape, generator=g)
  W = torch.randn((2 * ww, n_out), generator=g)
  want = torch.einsum("ij,j...->i...", W.double(), snaps.double())
  want = torch.complex(want[:ww], want[ww:])
  got = fdtdz_jax.project(snaps.cuda(), W.numpy())
  assert got.shape == want.shape and got.dtype == torch.complex64
  assert float((got.cpu() - want).abs().max()) <= 1e-6 * float(want.abs().max()) * n_out


def test_project_snapshots_uses_the_kernel_on_cuda_and_agrees_with_the_cpu_path():
  from pjz_b200 import _field as glue
  omega = np.array([2 * np.pi / 37, 2 * np.pi / 33, 2 * np.pi / 41])
  steps = (100, 100 + 7 * 9, 9)
  snaps = torch.randn((7, 3, 10, 12, 8), generator=torch.Generator().manual_seed(3))
  cpu = glue.project_snapshots(snaps, omega, steps, 0.5)
  gpu = glue.project_snapshots(snaps.cuda(), omega, steps, 0.5)
  assert gpu.is_cuda and gpu.dtype == torch.complex64
  torch.testing.assert_close(gpu.cpu(), cpu, rtol=1e-5, atol=1e-6 * float(cpu.abs().max()))


@pytest.mark.parametrize("nports", [1, 2, 3])
def test_fused_overlaps_match_the_eager_formula(nports):
  """b200fdtd_overlaps + host pinv vs the per-port slice-multiply-sum chains of `_overlap`
  (/root/reference/src/pjz/_field.py:305-338) for x, y and z ports, forward / backward / None."""
  from pjz_b200 import _field as glue
  g = torch.Generator().manual_seed(10 + nports)
  ww, xx, yy, zz = 2, 14, 12, 10
  fields = [torch.complex(torch.randn((ww, 3, xx, yy, zz), generator=g),
                          torch.randn((ww, 3, xx, yy, zz), generator=g)).cuda() for _ in range(nports)]
  shapes = [(ww, 2, 1, yy, zz), (ww, 2, xx, 1, zz), (ww, 2, xx, yy, 1)]
  modes = [torch.randn(shapes[i % 3], generator=g) for i in range(nports)]
  if nports > 2:
    modes[2] = torch.complex(modes[2], torch.randn(shapes[2], generator=g))
  betas = [np.array([0.31 + 0.02 * i, 0.35]) for i in range(nports)]
  pos = [4, 6, 5][:nports]
  is_fwd = [True, False, None][:nports]
  amps, svals = glue._overlaps_fused(fields, modes, betas, pos, is_fwd)
  for i in range(nports):
    want_a = (glue._overlap(modes[i], betas[i], pos[i], is_fwd[i], fields[i])[:, 0]
              if is_fwd[i] is not None else torch.ones(ww, dtype=torch.complex64).cuda())
    torch.testing.assert_close(amps[i], want_a, rtol=2e-5, atol=1e-5)
    for j in range(nports):
      want = glue._overlap(modes[j], betas[j], pos[j], is_fwd[j], fields[i])[:, 1] / want_a
      torch.testing.assert_close(svals[i][j], want, rtol=2e-4, atol=1e-5 * float(want.abs().max() + 1))
