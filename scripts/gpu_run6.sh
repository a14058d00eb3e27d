cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export SONIC_SPLITS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm_o python scripts/bench_gemm_one.py o 16 6 > gpurun_out/prof_gemm_o.log 2>&1
echo rc=$?; tail -3 gpurun_out/prof_gemm_o.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f -o gpurun_out/prof_gemm_gu python scripts/bench_gemm_one.py gateup 16 6 > gpurun_out/prof_gemm_gu.log 2>&1
echo rc=$?; tail -3 gpurun_out/prof_gemm_gu.log
ls -la gpurun_out/*.ncu-rep
