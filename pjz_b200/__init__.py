"""pjz_b200 -- B200-native FDTD engine behind pjz's ``fdtdz_jax.fdtdz`` entry point.

Public surface mirrors the part of ``pjz`` that sits on / next to the hot path
(/root/reference/src/pjz/__init__.py:3-16): ``field``, ``scatter``, ``SimParams``, ``mode``
(host harness) / ``mode_gpu`` (device-resident, batched over frequencies); ``decomposed_engine``
builds the multi-GPU, domain-decomposed ``engine=`` of ``field`` / ``scatter`` (no reference counterpart);
the engine itself is ``pjz_b200.fdtdz_jax.fdtdz`` (drop-in for the ``fdtdz_jax`` module pjz
imports at /root/reference/src/pjz/_field.py:6).
"""

from ._field import SimParams, field, scatter
from ._decomp import decomposed_engine
from ._mode import mode
from ._mode_gpu import mode_gpu

__all__ = ["SimParams", "field", "scatter", "mode", "mode_gpu", "decomposed_engine"]
