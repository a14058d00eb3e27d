cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,compute_mode,mig.mode.current --format=csv
nvidia-smi -q | grep -i -E "MPS|Compute Mode|SM " | head
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "bf16_path or batch_invariance or tensor_core_path or persistent" > gpurun_out/t16.log 2>&1; echo "== tests rc=$?"; tail -8 gpurun_out/t16.log
