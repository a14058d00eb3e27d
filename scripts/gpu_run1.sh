cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { # name timeout args...
  name=$1; to=$2; shift 2
  timeout $to python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short "$@" > gpurun_out/$name.log 2>&1
  echo "== $name rc=$?"; tail -15 gpurun_out/$name.log
}
run t1_mel 180 -k "mel_parity or mel_batch"
run t2_fp32 400 -k "fp32_path or staged or error_paths"
run t4_bf16 300 -k "bf16_path or batch_invariance or time_major"
