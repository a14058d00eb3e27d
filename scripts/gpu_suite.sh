# usage: gpu_suite.sh <tag> [bench batch sizes...]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "not full_model" > gpurun_out/${TAG}_tests.log 2>&1; echo "== tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
for B in "$@"; do
timeout 600 python bench.py --steps 3 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/${TAG}_b$B.json 2> gpurun_out/${TAG}_b$B.err
echo "== bench B=$B rc=$?"; tail -1 gpurun_out/${TAG}_b$B.err; python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_b$B.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','stage_ms_last_step','p50_latency_ms_single_20s_segment')}); print(d['profile_ms_by_class'])
PY
done
