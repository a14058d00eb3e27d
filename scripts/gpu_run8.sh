cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; to=$2; shift 2
  timeout $to python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short "$@" > gpurun_out/$name.log 2>&1
  echo "== $name rc=$?"; tail -6 gpurun_out/$name.log; }
run t11 600 -k "not full_model and not mel_parity and not gemm"
for B in 16 64; do
for NP in 0 1; do
SONIC_NO_PDL=$NP timeout 600 python bench.py --steps 2 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_pdl${NP}_b$B.json 2> gpurun_out/bench_pdl${NP}_b$B.err
echo "== bench B=$B NO_PDL=$NP rc=$?"; tail -1 gpurun_out/bench_pdl${NP}_b$B.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_pdl${NP}_b$B.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','stage_ms_last_step')})
PY
done; done
