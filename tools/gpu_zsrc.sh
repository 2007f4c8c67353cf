#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_decomp.py -m gpu -q -x -k "lean or large or long_run or y_session" > gpurun_out/pytest_zsrc.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_zsrc.log
B200FDTD_LEAN_TMA=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean_tilings or lean_ragged" > gpurun_out/pytest_zsrc_tma.log 2>&1; echo "pytest(TMA) rc=$?"; tail -2 gpurun_out/pytest_zsrc_tma.log
for g in 8 16 32; do timeout 600 python tools/bench_decomp.py --grid 1024 512 128 --tt 320 --mode y --ghost $g 2>&1 | tail -1 | cut -c1-420; done | tee gpurun_out/decomp_y_n1b.json
