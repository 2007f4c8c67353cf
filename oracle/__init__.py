"""CPU oracle for the FDTD hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as a fallback for the CUDA path.

PARITY UNPINNED at the engine boundary: the arithmetic this oracle restates
lives in the un-vendored PyPI dependency ``fdtdz>=1.1.3``
(/root/reference/setup.py:26; call site /root/reference/src/pjz/_field.py:254-269)
whose source is absent and un-fetchable, and no reference test exercises
``field()``/``scatter()``/``fdtdz_jax.fdtdz``.  The update equations are
therefore *defined* here (DESIGN.md section 3 lists every frozen decision);
what IS pinned are the reference's helper known-answer vectors (absorption
mask, PML sigma, sampling interval, ramped sine, mode betas), see
``tests/test_glue_kats.py``, and physics known-answers, see
``tests/test_oracle_physics.py``.
"""
