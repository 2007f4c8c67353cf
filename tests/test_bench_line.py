"""bench.py's JSON contract, exercised on the CPU: ``main()`` runs with the CUDA engine and the
``torch.cuda`` calls replaced by stand-ins (the C oracle plays the engine), so that a slip in the
line's assembly is caught here and not on the driver's GPU box.  Numbers mean nothing; keys do."""

import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
  spec = importlib.util.spec_from_file_location("bench_line_module", os.path.join(ROOT, "bench.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


class _Event:
  def __init__(self, enable_timing=False):
    pass

  def record(self):
    pass

  def elapsed_time(self, other):
    return 1.0


def _drop_device(fn):
  def wrapped(*a, **k):
    k.pop("device", None)
    return fn(*a, **k)
  return wrapped


@pytest.fixture
def fake_gpu(monkeypatch):
  from oracle import fdtd_c
  from pjz_b200 import fdtdz_jax

  def engine(**kw):
    as_torch = isinstance(kw["epsilon"], torch.Tensor)
    host = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    host["launch_params"] = None
    out = fdtd_c.fdtdz(**host)
    return torch.from_numpy(out) if as_torch else out

  monkeypatch.setattr(fdtdz_jax, "lib", lambda: None)
  monkeypatch.setattr(fdtdz_jax, "fdtdz", engine)
  monkeypatch.setattr(fdtdz_jax, "plan_info", lambda **kw: {
      "kernel": "systolic_lean", "tile_y": 15, "stages": 8, "threads": 288, "ctas": 144})
  monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
  monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
  monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
  monkeypatch.setattr(torch.cuda, "Event", _Event)
  monkeypatch.setattr(torch.cuda, "get_device_properties",
                      lambda d: types.SimpleNamespace(L2_cache_size=126 << 20))
  monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
  monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
  monkeypatch.setattr(torch, "empty", lambda *a, **k: torch.zeros(1, dtype=k.get("dtype", torch.float32)))
  monkeypatch.setattr(torch, "tensor", _drop_device(torch.tensor))
  for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
    monkeypatch.delenv(k, raising=False)


def _run(bench, argv, monkeypatch):
  monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
  buf = io.StringIO()
  with redirect_stdout(buf):
    bench.main()
  lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
  assert len(lines) == 1, "bench.py prints ONE JSON line"
  return json.loads(lines[0])


def test_both_arms_print_the_contract_keys_and_the_same_config(fake_gpu, monkeypatch):
  bench = _load_bench()
  common = ["--workload", "waveguide", "--tt", "40", "--steps", "2", "--warmup", "1"]
  line = _run(bench, common + ["--no-decomp", "--no-ab", "--no-reduced", "--cpu-seconds", "0.2"], monkeypatch)
  for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline",
            "cpu_baseline", "clocks", "plan"):
    assert k in line, k
  assert line["metric"] == "fdtd_cell_updates_per_s" and line["unit"] == "Gcell-updates/s"
  assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
  assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
  assert line["dtype"] == "f32" and line["data"] == "synthetic"
  assert "workload" in line["config"] and "model" not in line["config"] and "l2" in line["config"]
  assert line["gpu_launches"] > 0
  for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
    assert k in line["e2e"], k
  assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
  for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
    assert k in line["roofline"], k
  assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
  assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
  for k in ("value", "unit", "cores", "kind", "sample"):
    assert k in line["cpu_baseline"], k
  assert line["cpu_baseline"]["kind"] == "port"
  assert np.isfinite(line["value"]) and line["value"] > 0

  ref = _run(bench, common + ["--impl", "reference"], monkeypatch)
  assert ref["impl"] == "reference"
  for k in ("metric", "unit", "higher_is_better", "dtype", "config"):
    assert ref[k] == line[k], k                 # the driver compares the two arms on these
  assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
  assert ref["cpu_baseline"]["value"] == ref["value"] and ref["cpu_baseline"]["kind"] == "port"
