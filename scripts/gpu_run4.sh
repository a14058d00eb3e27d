cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2
  timeout $to python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short "$@" > gpurun_out/$name.log 2>&1
  echo "== $name rc=$?"; tail -12 gpurun_out/$name.log; }
run t8_splitk 200 -k "gemm"
run t9_bf16 300 -k "bf16 or batch_invariance"
timeout 600 python bench.py --steps 2 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
echo "== bench rc=$?"; tail -3 gpurun_out/bench3.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench3.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','stage_ms_last_step','gpu_launches')}); print(d['profile_ms_by_class']); print(d['roofline'])
PY
