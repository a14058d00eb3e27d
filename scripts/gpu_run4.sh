#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_sliced or pinned" > $O/pytest_rs.log 2>&1; echo "rs tests rc=$?"; tail -6 $O/pytest_rs.log
SONIC_DECODE_RS=1 timeout 200 python scripts/rs_phases.py bf16 1 2 4 8 16 > $O/rs_dbg.txt 2>&1; tail -24 $O/rs_dbg.txt
SONIC_DECODE_RS=1 timeout 200 python scripts/rs_phases.py int8 1 16 > $O/rs_dbg_int8.txt 2>&1; tail -16 $O/rs_dbg_int8.txt
