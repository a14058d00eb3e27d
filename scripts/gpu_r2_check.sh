#!/bin/bash
# full GPU suite + log-mel sweep (features only / with the time-major copy) + ncu capture of the frames kernel
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/pytest.log
timeout 300 python scripts/bench_mel.py 1024 > $O/mel_sweep.json 2>/dev/null; echo "mel rc=$?"; grep "^{'segments': \(1\|64\|1024\)," $O/mel_sweep.json
MEL_WITH_TM=1 MEL_ONLY_B=1024 timeout 300 python scripts/bench_mel.py 1024 2>/dev/null | grep "^{'segments"
SONIC_MEL_GROUP=1024 MEL_ONLY_B=1024 timeout 300 python scripts/bench_mel.py 1024 2>/dev/null | grep "^{'segments"
MEL_ONLY_B=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mel_frames_kernel -s 3 -c 1 -o $O/prof_mel_frames -f python scripts/bench_mel.py 64 > $O/ncu_mel.log 2>&1; echo "ncu mel rc=$?"
MEL_ONLY_B=64 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 12 -c 4 --csv --log-file $O/mel_launches_b64.csv python scripts/bench_mel.py 64 > /dev/null 2>&1; echo "ncu list rc=$?"
