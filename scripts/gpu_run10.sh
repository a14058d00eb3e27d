cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "int8" > gpurun_out/t13.log 2>&1; echo "== int8 tests rc=$?"; tail -15 gpurun_out/t13.log
for MODE in int8 bf16; do
timeout 600 python bench.py --steps 2 --warmup 3 --batch 16 --mode $MODE --no-cpu-baseline > gpurun_out/bench5_$MODE.json 2> gpurun_out/bench5_$MODE.err
echo "== bench $MODE rc=$?"; tail -1 gpurun_out/bench5_$MODE.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench5_$MODE.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','stage_ms_last_step')}); print(d['profile_ms_by_class'])
PY
done
