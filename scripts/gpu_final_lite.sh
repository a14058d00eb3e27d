# re-measure after a decode-kernel change: GPU tests, ncu traffic capture of the dominant kernel, default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/final_tests.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_persist_kernel -s 3 -c 1 -f -o gpurun_out/final_prof_persist_b64 python scripts/prof_persist.py 64 28 6 > gpurun_out/final_prof_persist.log 2>&1; echo "== ncu full rc=$?"
python scripts/ncu_traffic.py gpurun_out/final_prof_persist_b64.ncu-rep 64 bf16 gpurun_out/final_traffic > /dev/null && cp gpurun_out/final_traffic.json profiles/r01_traffic.json
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "== bench rc=$?"; tail -2 gpurun_out/final_bench.err
timeout 200 python bench.py --batch 16 --no-cpu-baseline > gpurun_out/final_bench_b16.json 2>/dev/null; echo "== bench b16 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/final_bench.json', 'gpurun_out/final_bench_b16.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'stage_ms_last_step', 'p50_latency_ms_single_20s_segment')})
    print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'traffic', 'avg_launch_ms')}, 'e2e', d['e2e']['value'])
PY
