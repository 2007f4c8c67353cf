#!/bin/bash
mkdir -p gpurun_out
B200FDTD_LEAN_TMA=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "lean_tilings or lean_ragged or large or long_run or fused_projection" > gpurun_out/pytest_tma_lean.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_tma_lean.log
cat > /tmp/v.txt <<'EOV'
--kernel systolic_lean
B200FDTD_LEAN_TMA=1 --kernel systolic_lean
--kernel systolic_lean
B200FDTD_LEAN_TMA=1 --kernel systolic_lean
B200FDTD_LEAN_UNROLL=2 --kernel systolic_lean
B200FDTD_LEAN_TMA=1 --kernel systolic_lean --workload demux
--kernel systolic_lean --workload demux
EOV
VARFILE=/tmp/v.txt TT=4000 STEPS=3 tools/gpu_misc.sh
