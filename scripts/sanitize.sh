#!/bin/bash
# compute-sanitizer over the hot path: memcheck (out-of-bounds / misaligned global, shared, TMA) and racecheck
# (shared-memory hazards between the warp roles) on the tiny model.  Run on a B200: scripts/sanitize.sh [outdir]
# Each tool/case pair runs under its own timeout; the summary lists errors per case.
OUT="${1:-gpurun_out/sanitize}"
mkdir -p "$OUT"
CS="${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}"
SUMMARY="$OUT/summary.txt"
: > "$SUMMARY"
for tool in ${SANITIZE_TOOLS:-memcheck racecheck}; do
  for c in ${SANITIZE_CASES:-smoke_fp32 smoke_bf16 b3 b20 b64 int8_b2 int8_b20}; do
    log="$OUT/${tool}_${c}.log"
    SANITIZE_TOKENS=3 timeout "${SANITIZE_TIMEOUT:-420}" "$CS" --tool "$tool" --error-exitcode 7 --print-limit 20 \
        python scripts/sanitize_target.py "$c" > "$log" 2>&1
    rc=$?
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $c rc=$rc ${errs:-no summary line (timeout or crash)}" | tee -a "$SUMMARY"
  done
done
