cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_tiny.csv python scripts/prof_tiny.py 16 6 2 > gpurun_out/prof_tiny.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_tiny.log; wc -l gpurun_out/launches_tiny.csv
