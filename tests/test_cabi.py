"""The C-ABI library: loads, exports every symbol include/b200fdtd.h declares, the header is
valid C and agrees with the ctypes mirror, and descriptor validation reports errors the way the
header promises.  No compute call is made (no GPU needed)."""

import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from pjz_b200 import fdtdz_jax
from tests.problems import random_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200fdtd.h")


def _declared_functions():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(b200fdtd_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
  names = _declared_functions()
  assert {"b200fdtd_run", "b200fdtd_run_host", "b200fdtd_workspace_bytes",
          "b200fdtd_xla_custom_call", "b200fdtd_last_error"} <= set(names)
  lib = ctypes.CDLL(fdtdz_jax.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/b200fdtd.h but not exported"


def test_header_is_plain_c_and_matches_ctypes_layout(built):
  with tempfile.TemporaryDirectory() as tmp:
    c = os.path.join(tmp, "t.c")
    open(c, "w").write(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b200fdtd.h"\n'
        'int main(void){printf("%zu %zu %zu %d\\n", sizeof(b200fdtd_desc),'
        ' offsetof(b200fdtd_desc, dt), offsetof(b200fdtd_desc, kernel), B200FDTD_ABI_VERSION);'
        'return 0;}\n')
    exe = os.path.join(tmp, "t")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "include"), c, "-o", exe])
    size, off_dt, off_kernel, abi = (int(v) for v in subprocess.check_output([exe]).split())
  assert size == ctypes.sizeof(fdtdz_jax.Desc)
  assert off_dt == fdtdz_jax.Desc.dt.offset
  assert off_kernel == fdtdz_jax.Desc.kernel.offset
  assert abi == fdtdz_jax.ABI_VERSION == fdtdz_jax.lib().b200fdtd_abi_version()


def test_validate_accepts_good_descriptor(built):
  d = fdtdz_jax.make_desc(**random_problem())
  assert fdtdz_jax.lib().b200fdtd_validate(ctypes.byref(d)) == 0
  assert (d.X, d.Y, d.Z) == (12, 10, 16)


@pytest.mark.parametrize("mutate,msg", [
    (lambda d: setattr(d, "struct_bytes", 3), "ABI"),
    (lambda d: setattr(d, "off_x", 99), "does not fit"),
    (lambda d: setattr(d, "source_position", -1), "source_position"),
    (lambda d: setattr(d, "pml_lo", 1000), "pml_widths"),
    (lambda d: setattr(d, "out_step", 0), "output_steps"),
    (lambda d: setattr(d, "dt", 0.0), "dt"),
    (lambda d: setattr(d, "kernel", 7), "kernel"),
])
def test_validate_reports_errors(built, mutate, msg):
  d = fdtdz_jax.make_desc(**random_problem())
  mutate(d)
  lib = fdtdz_jax.lib()
  assert lib.b200fdtd_validate(ctypes.byref(d)) == 1  # B200FDTD_EINVAL
  assert msg in lib.b200fdtd_last_error().decode()
  assert lib.b200fdtd_num_outputs(ctypes.byref(d)) == -1
  assert lib.b200fdtd_output_bytes(ctypes.byref(d)) == 0


def test_wrapper_shape_errors(built):
  kw = random_problem()
  bad = dict(kw); bad["epsilon"] = kw["epsilon"][:2]
  with pytest.raises(ValueError, match="epsilon"):
    fdtdz_jax.make_desc(**bad)
  bad = dict(kw); bad["source_field"] = kw["source_field"][..., :-1]
  with pytest.raises(ValueError, match="source_field"):
    fdtdz_jax.make_desc(**bad)
  bad = dict(kw); bad["pml_sigma"] = kw["pml_sigma"][:-1]
  with pytest.raises(ValueError, match="pml"):
    fdtdz_jax.make_desc(**bad)
  bad = dict(kw); bad["launch_params"] = {"kernel": "warp9"}
  with pytest.raises(ValueError, match="kernel"):
    fdtdz_jax.make_desc(**bad)
  bad = dict(kw); bad["output_steps"] = (0, 1000, 1)
  with pytest.raises(ValueError, match="output_steps"):
    fdtdz_jax.make_desc(**bad)


@pytest.mark.parametrize("steps,ok", [((2, 12, 3), False), ((3, 13, 4), False), ((2, 11, 3), True),
                                      ((9, 10, 1), True), ((10, 11, 1), False), ((0, 10, 1), True)])
def test_last_snapshot_must_be_a_step_of_the_run(built, steps, ok):
  """tt = 10: a snapshot at step >= tt would never be written ((2,12,3) -> [2,5,8,11]
  used to be accepted); the NumPy oracle rejects the same inputs."""
  from oracle import fdtd_numpy
  kw = random_problem(tt=10, output_steps=steps)
  lib = fdtdz_jax.lib()
  d = fdtdz_jax.make_desc(**{**kw, "output_steps": (2, 9, 3)})
  d.out_start, d.out_stop, d.out_step = steps
  assert (lib.b200fdtd_validate(ctypes.byref(d)) == 0) == ok
  if not ok:
    assert "output_steps" in lib.b200fdtd_last_error().decode()
    with pytest.raises(ValueError):
      fdtd_numpy.fdtdz(**kw)


def test_removed_kernel_value_is_rejected(built):
  d = fdtdz_jax.make_desc(**random_problem())
  d.kernel = 4                                  # was the TMA-staged variant (removed in round 2)
  assert fdtdz_jax.lib().b200fdtd_validate(ctypes.byref(d)) == 1


def test_output_count(built):
  kw = random_problem(tt=30, output_steps=(4, 30, 7))
  d = fdtdz_jax.make_desc(**kw)
  lib = fdtdz_jax.lib()
  assert lib.b200fdtd_num_outputs(ctypes.byref(d)) == len(range(4, 30, 7))
  assert lib.b200fdtd_output_bytes(ctypes.byref(d)) == 4 * 3 * d.xx * d.yy * d.zz * 4


def test_no_cpu_fallback(built):
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(RuntimeError, match="CUDA"):
    fdtdz_jax.fdtdz(**random_problem())


def test_install_registers_module_name():
  import sys
  fdtdz_jax.install()
  import fdtdz_jax as shim  # the name pjz imports (/root/reference/src/pjz/_field.py:6)
  assert shim.fdtdz is fdtdz_jax.fdtdz
  del sys.modules["fdtdz_jax"]


def test_foreign_device_arrays_are_adopted_through_dlpack():
  """Anything that speaks DLPack and is neither torch nor NumPy (JAX, CuPy) enters zero-copy;
  torch tensors, NumPy arrays, lists and None pass through untouched."""
  import torch
  from pjz_b200 import fdtdz_jax

  class Foreign:                                  # stands in for a jax.Array
    def __init__(self, t):
      self._t = t
      self.shape = tuple(t.shape)

    def __dlpack__(self, *a, **k):
      return self._t.__dlpack__(*a, **k)

    def __dlpack_device__(self):
      return self._t.__dlpack_device__()

  t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
  got, was = fdtdz_jax._adopt(Foreign(t))
  assert was and isinstance(got, torch.Tensor) and got.data_ptr() == t.data_ptr()
  assert torch.equal(got, t)
  for same in (t, t.numpy(), [1.0, 2.0], None):
    got, was = fdtdz_jax._adopt(same)
    assert got is same and not was

  class Broken:
    shape = (1,)

    def __dlpack__(self, *a, **k):
      raise TypeError("no")

    def __dlpack_device__(self):
      return (1, 0)
  b = Broken()
  got, was = fdtdz_jax._adopt(b)
  assert got is b and not was
  assert fdtdz_jax._return_as("cupy", t) is t
