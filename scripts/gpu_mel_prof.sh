#!/bin/bash
# mel parity, per-launch times of one 1024-segment sweep point, sweep tail
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "mel" 2>&1 | tail -2
MEL_ONLY_B=1024 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 40 --csv --log-file $O/mel_launches.csv python scripts/bench_mel.py 1024 > /dev/null 2>&1; echo "ncu list rc=$?"
MEL_ONLY_B=256 timeout 200 python scripts/bench_mel.py 1024 2>&1 | grep "^{'segments" | head -8
SONIC_MEL_GROUP=1024 MEL_ONLY_B=1024 timeout 200 python scripts/bench_mel.py 1024 2>&1 | grep "^{'segments" | head -8
