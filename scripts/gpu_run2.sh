#!/bin/bash
# round-2 GPU run 2: first contact of the row-sliced decode kernel
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "row_sliced" > $O/pytest_rs.log 2>&1; echo "rs tests rc=$?"; tail -25 $O/pytest_rs.log
timeout 200 python scripts/rs_phases.py bf16 1 4 16 32 > $O/rs_phases_bf16.txt 2>&1; echo "phases rc=$?"; tail -12 $O/rs_phases_bf16.txt
timeout 200 python scripts/rs_phases.py int8 1 16 > $O/rs_phases_int8.txt 2>&1; echo "phases int8 rc=$?"; tail -6 $O/rs_phases_int8.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_all.log 2>&1; echo "all tests rc=$?"; tail -15 $O/pytest_all.log
timeout 300 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bf16_b1.json 2> $O/bench_bf16_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --batch 16 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_bf16_b16.json 2> $O/bench_bf16_b16.err; echo "b16 rc=$?"
timeout 300 python bench.py --mode int8 --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_int8_b1.json 2> $O/bench_int8_b1.err; echo "int8 rc=$?"
timeout 300 python bench.py --workload realtime --steps 3 --warmup 3 > $O/bench_realtime.json 2> $O/bench_realtime.err; echo "realtime rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"],1), d.get("decode",{}).get("ms_per_token_step"), (d.get("roofline") or {}).get("frac"), d.get("interim_ms",{}).get("p50"), d.get("committed_ms",{}).get("p50"))
    except Exception as e:
        print(f, "ERR", e)
PY
