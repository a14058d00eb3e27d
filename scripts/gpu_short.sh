#!/bin/bash
# short-window (streaming) encoder: parity tests, then the realtime workload with it
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_asr_model.py -q -x -k "short_window" 2>&1 | tail -4
SONIC_SHORT_WINDOW_MAX_NEW=15 timeout 150 python bench.py --workload realtime --steps 2 2>/dev/null > gpurun_out/realtime_short_wm.json; python -c "
import json; d=json.loads(open('gpurun_out/realtime_short_wm.json').read().strip().splitlines()[-1]); print('short window', d['interim_ms'], d['committed_ms']['p50'])"
