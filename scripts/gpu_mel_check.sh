#!/bin/bash
# mel parity tests + the config-5 sweep + the int8 / bf16 batch-1 decode step of the current build
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "mel" 2>&1 | tail -5
MEL_ONLY_B=${MEL_B:-64} timeout 200 python scripts/bench_mel.py ${MEL_MAX:-1024} 2>&1 | grep "segments" | head -8
for m in int8 bf16; do
timeout 200 python bench.py --batch 1 --mode $m --no-cpu-baseline --no-api-threads --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HEAD $m', round(d['decode']['ms_per_token_step'],4), d['stage_ms_last_step'])"
done
