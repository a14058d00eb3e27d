cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_persist_kernel -s 2 -c 1 -f -o gpurun_out/prof_persist_b64 python scripts/prof_persist.py 64 4 5 > gpurun_out/prof_persist_b64.log 2>&1
echo rc=$?; tail -3 gpurun_out/prof_persist_b64.log
