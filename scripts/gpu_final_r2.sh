#!/bin/bash
# round-2 final measurements: full GPU suite, ncu capture of the dominant kernel (traffic), bench lines, launch list, mel sweep
cd "$(dirname "$0")/.."
O=gpurun_out/final2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --tb=short > $O/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $O/pytest.log
MP=320 MN=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_persist_kernel -s 3 -c 1 -f -o $O/prof_decode_b256 python scripts/prof_persist.py 256 28 6 > $O/prof_decode.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_traffic.py $O/prof_decode_b256.ncu-rep 256 bf16 $O/traffic > /dev/null 2>&1 && cp $O/traffic.json profiles/r02_traffic.json && cp $O/traffic.csv profiles/r02_ncu_full_decode_b256.csv
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -1 $O/bench_default.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline > $O/bench_bf16_b1.json 2>/dev/null; echo "b1 rc=$?"
timeout 300 python bench.py --batch 1 --mode int8 --no-cpu-baseline > $O/bench_int8_b1.json 2>/dev/null; echo "int8 b1 rc=$?"
timeout 400 python bench.py --mode int8 --no-cpu-baseline --no-api-threads --steps 3 > $O/bench_int8_b128.json 2>/dev/null; echo "int8 b128 rc=$?"
timeout 300 python bench.py --workload realtime > $O/bench_realtime.json 2>/dev/null; echo "realtime rc=$?"
timeout 300 python bench.py --workload file1h --steps 2 --warmup 1 > $O/bench_file1h_1gpu.json 2>/dev/null; echo "file1h rc=$?"
timeout 300 python scripts/bench_mel.py 1024 > $O/mel_sweep.json 2>/dev/null; echo "mel rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api-threads > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final2/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f.split('/')[-1], {k: d.get(k) for k in ('value', 'ms_per_step', 'stage_ms_last_step', 'p50_latency_ms_single_20s_segment')})
    if d.get('roofline'): print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'traffic', 'avg_launch_ms', 'bytes_per_launch', 'share_of_step')})
    if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'))
    if d.get('e2e'): print('   e2e', d['e2e'].get('value'))
    if d.get('api_threads_asrmodel_transcribe'): print('   api threads', d['api_threads_asrmodel_transcribe'])
    if d.get('interim_ms'): print('   interim', d['interim_ms'].get('p50'), d['interim_ms'].get('p95'), 'committed', d['committed_ms'].get('p50'))
PY
