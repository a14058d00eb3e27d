#!/bin/bash
# session 2 baseline: full GPU suite, int8 B=64 with the expand-to-bf16 prefill/encoder path, default bench
cd "$(dirname "$0")/.."
O=gpurun_out/s2a; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -8 $O/pytest.log
timeout 400 python bench.py --batch 64 --mode int8 --no-cpu-baseline --no-api-threads --steps 3 > $O/bench_int8_b64.json 2> $O/bench_int8_b64.err; echo "int8 b64 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s2a/bench_int8_b64.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','stage_ms_last_step')}); print(d['profile_ms_by_class'])
PY
