#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel" > $O/pytest_mel.log 2>&1; echo "mel tests rc=$?"; tail -4 $O/pytest_mel.log
timeout 300 python scripts/bench_mel.py 1024 > $O/mel_sweep.json 2> $O/mel_sweep.err; echo "sweep rc=$?"; grep -E "segments" $O/mel_sweep.json | tail -11 | cut -c1-140
export SONIC_DECODE_RS=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_sliced" > $O/pytest_rs.log 2>&1; echo "rs tests rc=$?"; tail -3 $O/pytest_rs.log
for pf in 0 1; do echo "L2_PREFETCH=$pf"; SONIC_RS_L2_PREFETCH=$pf timeout 200 python scripts/rs_phases.py bf16 1 16 2>&1 | grep -E "step [0-9]|dbg gate|dbg down"; SONIC_RS_L2_PREFETCH=$pf timeout 200 python scripts/rs_phases.py int8 1 2>&1 | grep -E "step [0-9]|dbg gate"; done
