# final round measurements: full GPU test suite, bench lines, ncu launch list + full capture of the dominant kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/final_tests.log 2>&1; echo "== tests rc=$?"; tail -3 gpurun_out/final_tests.log
# traffic of the dominant kernel first, so that the bench line can quote it
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_persist_kernel -s 3 -c 1 -f -o gpurun_out/final_prof_persist_b64 python scripts/prof_persist.py 64 28 6 > gpurun_out/final_prof_persist.log 2>&1; echo "== ncu full rc=$?"
python scripts/ncu_traffic.py gpurun_out/final_prof_persist_b64.ncu-rep 64 bf16 gpurun_out/final_traffic; cp gpurun_out/final_traffic.json profiles/r01_traffic.json
ls -la gpurun_out/final_prof_persist_b64.ncu-rep
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "== bench rc=$?"; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "== ref rc=$?"
for b in 1 16 32; do timeout 300 python bench.py --batch $b --no-cpu-baseline > gpurun_out/final_bench_b$b.json 2>/dev/null; echo "== bench b$b rc=$?"; done
timeout 300 python bench.py --batch 1 --mode int8 --no-cpu-baseline > gpurun_out/final_bench_int8_b1.json 2>/dev/null; echo "== bench int8 rc=$?"
timeout 300 python bench.py --batch 64 --mode int8 --no-cpu-baseline > gpurun_out/final_bench_int8_b64.json 2>/dev/null; echo "== bench int8 b64 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1; echo "== ncu list rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'stage_ms_last_step', 'p50_latency_ms_single_20s_segment', 'gpu_launches')})
    if 'roofline' in d and d['roofline']: print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'traffic', 'avg_launch_ms', 'bytes_per_launch')})
    if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'))
    if d.get('e2e'): print('   e2e', d['e2e'].get('value'))
PY
