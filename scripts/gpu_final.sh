# final round measurements: full GPU test suite, bench line, ncu launch list + full capture of the dominant kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/final_tests.log 2>&1; echo "== tests rc=$?"; tail -4 gpurun_out/final_tests.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "== bench rc=$?"; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "== ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --batch 8 --max-new 8 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1; echo "== ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_persist_kernel -s 3 -c 1 -f -o gpurun_out/final_prof_persist_b16 python scripts/prof_persist.py 16 28 6 > gpurun_out/final_prof_persist.log 2>&1; echo "== ncu full rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','stage_ms_last_step','p50_latency_ms_single_20s_segment','gpu_launches','clocks')})
print(d['roofline']); print(d['cpu_baseline']); print(d['e2e'])
PY
