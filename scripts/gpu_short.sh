#!/bin/bash
# short-window (streaming) encoder: parity tests + the realtime workload with and without it
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_asr_model.py -q -x -k "short_window or features_only" 2>&1 | tail -15
timeout 300 python bench.py --workload realtime --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full window', d['interim_ms'], d['committed_ms']['p50'])"
SONIC_SHORT_WINDOW_MAX_NEW=15 timeout 300 python bench.py --workload realtime --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('short window', d['interim_ms'], d['committed_ms']['p50'])"
