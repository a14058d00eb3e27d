cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mel_frames_kernel -s 3 -c 1 -f -o gpurun_out/prof_mel python scripts/bench_mel.py 64 > gpurun_out/prof_mel.log 2>&1
echo rc=$?
