#!/bin/bash
# round-2 closing measurements at the final commit: full GPU suite, smoke, mel sweep, bench lines of every BASELINE config,
# launch list of one default step, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out/final3; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python scripts/bench_mel.py 1024 > $O/mel_sweep.json 2>/dev/null; echo "mel rc=$?"; grep "^{'segments': \(64\|1024\)," $O/mel_sweep.json
MEL_WITH_TM=1 MEL_ONLY_B=1024 timeout 300 python scripts/bench_mel.py 1024 2>/dev/null | grep "^{'segments" > $O/mel_with_tm_1024.txt; cat $O/mel_with_tm_1024.txt
SONIC_NO_PDL=1 MEL_ONLY_B=1024 timeout 300 python scripts/bench_mel.py 1024 2>/dev/null | grep "^{'segments" > $O/mel_no_pdl_1024.txt; cat $O/mel_no_pdl_1024.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -1 $O/bench_default.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline > $O/bench_bf16_b1.json 2>/dev/null; echo "b1 rc=$?"
timeout 300 python bench.py --batch 1 --mode int8 --no-cpu-baseline > $O/bench_int8_b1.json 2>/dev/null; echo "int8 b1 rc=$?"
timeout 400 python bench.py --mode int8 --no-cpu-baseline --no-api-threads --steps 3 > $O/bench_int8_b128.json 2>/dev/null; echo "int8 b128 rc=$?"
timeout 300 python bench.py --workload realtime > $O/bench_realtime.json 2>/dev/null; echo "realtime rc=$?"
SONIC_SHORT_WINDOW_MAX_NEW=15 timeout 300 python bench.py --workload realtime > $O/bench_realtime_short_window.json 2>/dev/null; echo "realtime short rc=$?"
timeout 300 python bench.py --workload file1h --steps 2 --warmup 1 > $O/bench_file1h_1gpu.json 2>/dev/null; echo "file1h rc=$?"
# -c 4400: set-up + the warm-up / timed 256-segment steps (595 launches each); the single-segment latency probes that follow are not needed
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api-threads > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?"
SONIC_REF_BUDGET_S=60 timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final3/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f.split('/')[-1], {k: d.get(k) for k in ('value', 'ms_per_step', 'stage_ms_last_step', 'p50_latency_ms_single_20s_segment')})
    if d.get('roofline'): print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'traffic', 'avg_launch_ms', 'bytes_per_launch', 'share_of_step')})
    if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'), d['cpu_baseline'].get('kind'))
    if d.get('e2e'): print('   e2e', d['e2e'].get('value'))
    if d.get('api_threads_asrmodel_transcribe'): print('   api threads', d['api_threads_asrmodel_transcribe'])
    if d.get('interim_ms'): print('   interim', d['interim_ms'].get('p50'), d['interim_ms'].get('p95'), 'committed', d['committed_ms'].get('p50'))
PY
