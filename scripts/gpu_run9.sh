cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/bench_gemm.py 2>&1 | head -10
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "gemm or bf16 or batch_inv" > gpurun_out/t12.log 2>&1; echo "== tests rc=$?"; tail -4 gpurun_out/t12.log
for B in 16 64; do
timeout 600 python bench.py --steps 2 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench4_b$B.json 2> gpurun_out/bench4_b$B.err
echo "== bench B=$B rc=$?"; tail -1 gpurun_out/bench4_b$B.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench4_b$B.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','stage_ms_last_step')})
PY
done
