#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
export SONIC_DECODE_RS=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_sliced" > $O/pytest_rs.log 2>&1; echo "rs tests rc=$?"; tail -3 $O/pytest_rs.log
for pf in 0 1; do echo "L2_PREFETCH=$pf"; SONIC_RS_L2_PREFETCH=$pf timeout 200 python scripts/rs_phases.py bf16 1 16 2>&1 | grep -E "step [0-9]|dbg gate|dbg down"; SONIC_RS_L2_PREFETCH=$pf timeout 200 python scripts/rs_phases.py int8 1 2>&1 | grep -E "step [0-9]|dbg gate"; done
