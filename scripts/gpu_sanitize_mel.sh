#!/bin/bash
# compute-sanitizer memcheck / racecheck over the rewritten log-mel kernels (scripts/sanitize_target.py mel)
O=gpurun_out/sanitize_mel; mkdir -p $O
for tool in ${SANITIZE_TOOLS:-memcheck racecheck}; do
  timeout 100 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scripts/sanitize_target.py mel > $O/${tool}_mel.log 2>&1
  echo "$tool mel rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${tool}_mel.log | tail -1)"; grep "sanitize-target" $O/${tool}_mel.log
done
