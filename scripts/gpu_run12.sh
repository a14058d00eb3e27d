cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "bf16_path or batch_invariance or tensor_core_path" > gpurun_out/t15.log 2>&1; echo "== tests rc=$?"; tail -12 gpurun_out/t15.log
timeout 200 python scripts/persist_phases.py 4 2>&1 | tail -6
