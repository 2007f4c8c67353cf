"""Regenerates tests/golden/engine_golden.npz.

The reference engine (PyPI fdtdz) cannot be imported here, so these vectors
are produced by the float64 NumPy spec ``oracle/fdtd_numpy.py`` and frozen: they pin the
spec itself (any later edit of the oracle that changes results fails test_oracle.py), the C
restatement, and the CUDA engine to one committed set of numbers.
Run:  python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fdtd_numpy  # noqa: E402
from tests.problems import random_problem  # noqa: E402

CASES = {
    "x_src": dict(domain=(12, 10, 16), axis=0, pml=(3, 4), tt=24, seed=100, output_steps=(10, 24, 5)),
    "y_src": dict(domain=(10, 12, 12), axis=1, pml=(2, 2), tt=24, seed=101, output_steps=(11, 24, 4)),
    "z_src": dict(domain=(9, 11, 20), axis=2, pml=(4, 6), tt=24, seed=102, output_steps=(23, 24, 1)),
    "no_pml": dict(domain=(8, 8, 8), axis=0, pml=(0, 0), tt=16, seed=103, output_steps=(0, 16, 5)),
    "z_batch": dict(domain=(10, 9, 8), axis=0, pml=(0, 0), tt=20, seed=104, z_as_batch=True,
                    output_steps=(15, 20, 2)),
    "ragged_z": dict(domain=(7, 9, 13), axis=2, pml=(3, 5), tt=18, seed=105, output_steps=(9, 18, 3)),
    # column heights that select the warp-per-column-pair kernel (32 fp32 vectors) and its
    # sub-warp variants (16 and 8 vectors) under launch_params "auto" / "systolic_lean"
    "tall_128": dict(domain=(6, 9, 128), sub=(3, 4, 120), axis=0, pml=(16, 16), tt=20, seed=106,
                     output_steps=(13, 20, 6)),
    "short_64": dict(domain=(6, 9, 64), sub=(3, 4, 60), axis=1, pml=(8, 8), tt=20, seed=107,
                     output_steps=(13, 20, 6)),
    "short_30": dict(domain=(5, 11, 30), sub=(3, 5, 26), axis=2, pml=(5, 7), tt=20, seed=108,
                     output_steps=(13, 20, 6)),
}


def main():
  out = {}
  for name, spec in CASES.items():
    kw = random_problem(**spec)
    out[name] = fdtd_numpy.fdtdz(**kw)
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "engine_golden.npz")
  np.savez_compressed(path, **out)
  print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
  main()
