cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
timeout 900 python bench.py --steps 2 --warmup 3 --batch 16 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "== bench rc=$?"; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json | cut -c1-3000
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "full_model" > gpurun_out/t5_full.log 2>&1
echo "== full rc=$?"; tail -15 gpurun_out/t5_full.log
