#!/bin/bash
# decode kernel iteration: parity tests of the persistent classes, per-phase stamps, B=64 / B=32 bench lines
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-s2b}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or pinned or server_sized or batch_invariance or bf16_path" > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -4 $O/pytest.log
timeout 300 python scripts/persist_phases.py 4 2>&1 | grep -v Warning | tee $O/phases.txt
for B in ${2:-64}; do
timeout 400 python bench.py --batch $B --no-cpu-baseline --no-api-threads --steps 3 > $O/bench_b$B.json 2> $O/bench_b$B.err; echo "bench b$B rc=$?"
python - <<PY
import json
d = json.loads(open('$O/bench_b$B.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','stage_ms_last_step')}); print(d['decode'], d['roofline']['frac'])
PY
done
