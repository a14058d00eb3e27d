#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05 or pinned or graph_path" > $O/pytest_tc.log 2>&1; echo "tc tests rc=$?"; tail -4 $O/pytest_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['value'],1), d['ms_per_step'], d['decode'], d['roofline']['frac'])"
NCU="ncu --set full --clock-control none --import-source on"
MEL_ONLY_B=64 timeout 300 $NCU -k regex:mel_frames_kernel -s 3 -c 1 -o $O/prof_mel_frames -f python scripts/bench_mel.py 64 > $O/ncu_mel.log 2>&1; echo "ncu mel rc=$?"
export SONIC_BENCH_MINWARM=1
timeout 400 $NCU -k regex:gemm_tc_persist_kernel -s 40 -c 2 -o $O/prof_gemm_persist -f python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-api-threads > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 400 $NCU -k regex:attention_tc_kernel -s 8 -c 1 -o $O/prof_enc_attn -f python bench.py --batch 16 --steps 1 --warmup 1 --no-cpu-baseline --no-api-threads > $O/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 400 $NCU -k regex:decode_persist_kernel -s 20 -c 1 -o $O/prof_decode_b64 -f python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-api-threads > $O/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 573 -c 573 --csv --log-file $O/launches_b64_step.csv python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-api-threads > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la $O
