#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -14 $O/pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bf16_b1.json 2> $O/bench_bf16_b1.err
timeout 300 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bf16_b16.json 2> $O/bench_bf16_b16.err
timeout 300 python bench.py --mode int8 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_int8_b64.json 2> $O/bench_int8_b64.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "dec ms/tok", d.get("decode",{}).get("ms_per_token_step"), "frac", (d.get("roofline") or {}).get("frac"), d.get("stage_ms_last_step"))
        print("    ", d.get("profile_ms_by_class"))
    except Exception as e:
        print(f, "ERR", e)
PY
